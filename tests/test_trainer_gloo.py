"""Host-side logic of the data-parallel trainer on CPU with world_size 2 (gloo): flat parameter/gradient buffers,
LR grouping, batch sharding and the gradient / loss-statistics all-reduce plumbing.  The kernels themselves need a
GPU (tests/test_gpu_parity.py::test_trainer_direct_grad_accumulation_matches_autograd, bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from deer_b200.trainer import DEERDataParallelTrainer, FlatBuffers, reference_lr_group, shard_batch


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.audio_encoder = nn.Linear(5, 3)      # name contains "encoder" -> LR group 0 (training.py:128-142)
        self.head = nn.Linear(3, 2)


def test_flat_buffers_rehome_parameters_and_group_bounds():
    torch.manual_seed(0)
    m = Tiny()
    ref = {k: v.clone() for k, v in m.state_dict().items()}
    fb = FlatBuffers(m, reference_lr_group)
    assert fb.payload == sum(p.numel() for p in m.parameters())
    assert fb.numel % 4 == 0 and fb.numel >= fb.payload
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref[k])                                  # values preserved
    for p in m.parameters():
        assert p.data_ptr() >= fb.params.data_ptr() and p.grad is not None
        assert p.grad.data_ptr() >= fb.grads.data_ptr()
        assert p.data_ptr() % 16 == 0                                  # 16-byte aligned slots
    groups = [g for g, _, _ in fb.group_bounds]
    assert groups == [0, 1]
    (g0, lo0, hi0), (g1, lo1, hi1) = fb.group_bounds
    assert lo0 == 0 and hi0 == lo1 and hi1 == fb.numel
    assert all("encoder" in n for n in fb.names[:2]) and all("encoder" not in n for n in fb.names[2:])
    m.head.weight.grad.fill_(2.0)                                      # views alias the flat gradient buffer
    assert float(fb.grads.sum()) == 2.0 * m.head.weight.numel()


def test_shard_batch_is_contiguous_and_exhaustive():
    b = {"audio_features": torch.arange(24.).view(8, 3), "targets": torch.arange(8.).view(8, 1)}
    parts = [shard_batch(b, r, 4) for r in range(4)]
    assert all(p["targets"].shape[0] == 2 for p in parts)
    assert torch.equal(torch.cat([p["audio_features"] for p in parts]), b["audio_features"])


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                      # identical replicas
        tr = DEERDataParallelTrainer(Tiny())
        assert tr.world == world
        f = tr.flat
        f.grads.fill_(float(rank + 1))            # stand-in for the shard gradients written by the backward kernels
        tr._allreduce(f.grads)
        assert torch.allclose(f.grads, torch.full_like(f.grads, 3.0))          # 1 + 2: gradients are SUMMED
        stats = torch.full((3, 40), float(rank + 1))                           # DEER loss sufficient statistics
        tr._allreduce(stats)
        assert torch.allclose(stats, torch.full_like(stats, 3.0))
        g = torch.Generator().manual_seed(1)
        batch = {"x": torch.randn(6, 4, generator=g)}
        mine = shard_batch(batch, rank, world)["x"]
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert torch.equal(torch.cat(gathered), batch["x"])
    finally:
        dist.destroy_process_group()


def test_gradient_and_statistics_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
