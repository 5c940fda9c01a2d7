"""Parity at the BENCHMARKED dispatch (BASELINE.json configs 2-5, SURVEY.md section 8d "Parity gates").

The golden / oracle composites in test_gpu_parity.py run at B <= 7, where every contraction of the path is small; the
tests below run the default-policy path at the batch sizes bench.py times -- tcgen05 TF32 scorers / Conv1d taps /
projections, FP16/BF16 LSTM GEMMs and FP16 pre-activations, the sliding-window Conv1d, the two-stream encoder fork, the
dual-sub-tile inference LSTM kernel, CUDA-graph replay -- against the fp64 CPU oracle (oracle/deer_oracle.py,
golden-pinned by tests/test_oracle_golden.py) on the same seeded inputs and weights.

Gates (north_star): NIG parameters and every loss component <= 1e-3 relative; gradients cosine >= 0.999 per parameter
tensor and over the flat buffer; exactly-zero reference gradients exactly zero.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import assert_close, cosine, rel_l2

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import deer_b200
    from deer_b200 import _lib, ops
    from deer_b200.trainer import DEERDataParallelTrainer, capture_forward
    from gen_common import det_state_dict, pooled_inputs, seq_inputs
    from oracle import deer_oracle as O

DEV = "cuda"
TOL = 1e-3
TA, TV, TT = 300, 50, 64      # bench.py shapes
DIMS = ("valence", "arousal", "dominance")


def cu(t):
    return t.to(torch.float32).to(DEV)


@pytest.fixture(autouse=True)
def _default_policy():
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    ops.set_direct_grad_accumulation(False)
    yield
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    ops.set_direct_grad_accumulation(False)


def _seq_model(seed, dropout=0.0):
    torch.manual_seed(0)
    model = deer_b200.SequenceDEERModel(dropout=dropout)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd64 = det_state_dict(shapes, seed=seed)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd64.items()})
    return model.to(DEV), sd64


def _nig_cols(ref, key):
    return torch.cat([ref[f"{d}_{key}"] for d in DIMS], dim=1)


def _check_nig(out, ref, what):
    """NIG parameters of all three dimensions, <= 1e-3 relative l2 and <= 1e-3 of the tensor scale element-wise."""
    for ours, theirs in (("mu_all", ref["mu_all"]), ("nu", _nig_cols(ref, "nu")), ("alpha", _nig_cols(ref, "alpha")),
                         ("beta", _nig_cols(ref, "beta")), ("uncertainty_all", ref["uncertainty_all"])):
        got = out[ours].detach().double().cpu()
        assert_close(got, theirs.detach(), TOL, f"{what}:{ours}")
    for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features"):
        assert_close(out[k], ref[k].detach(), TOL, f"{what}:{k}")


def _check_grads(model, sdg):
    flat_c, flat_o = [], []
    worst = (1.0, None)
    for n, p in model.named_parameters():
        og = sdg[n].grad
        if og is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        if float(og.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, n          # MHA Q/K rows of the seq-1 attention: exactly zero
            continue
        if float(og.abs().max()) < 1e-12:   # exactly-zero derivative up to round-off (bias before a softmax, ...)
            assert float(p.grad.abs().max()) < 1e-5, n
            continue
        c = cosine(p.grad, og)
        if c < worst[0]:
            worst = (c, n)
        assert c >= 0.999, (n, c)
        flat_c.append(p.grad.flatten().double().cpu())
        flat_o.append(og.flatten())
    cf = cosine(torch.cat(flat_c), torch.cat(flat_o))
    assert cf >= 0.999, cf
    return worst, cf


def test_train_step_b256_benchmark_dispatch_vs_fp64_oracle():
    """BASELINE configs[2]: B=256 training step (parity variant, dropout 0), T = 300/50/64, ragged text masks,
    default precision policy, two-stream encoder fork -- forward outputs, all 17 loss components and every parameter
    gradient against the fp64 oracle.  Asserts that the tensor-core engines really ran."""
    B = 256
    model, sd64 = _seq_model(seed=21)
    model.train()
    batch = seq_inputs(B, TA, TV, TT, seed=21)
    before = _lib.engine_counts()
    out = model(*[cu(t) for t in batch[:5]])
    loss = model.compute_loss(out, cu(batch[5]))
    loss["total_loss"].backward()
    torch.cuda.synchronize()
    after = _lib.engine_counts()
    used = {k: after[k] - before[k] for k in after}
    # the benchmarked dispatch: split-precision 16-bit tcgen05 for the forward scorers / conv taps / projections, TF32
    # tcgen05 for their backward, 16-bit tcgen05 for the LSTM GEMMs, fused 3xTF32 for the post-pooling chain
    assert used["h16_split"] >= 5, used
    assert used["tf32_pair"] + used["tf32"] >= 8, used
    assert used["h16"] >= 8, used
    # (>= 40 separate fused 3xTF32 nodes module by module; with the persistent chain kernel the whole fusion + head is
    # one dispatch per direction, the encoders' small layers remain separate)
    from deer_b200 import chain
    assert used["tf32x3"] >= (10 if chain.enabled() else 40) and used["simt"] == 0, used
    assert ops.branch_streams_enabled() and B <= ops.branch_max_batch()

    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd64.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref, rloss = O.sequence_model_loss(*batch[:5], batch[5], sdg, training=True)
    rloss["total_loss"].backward()
    _check_nig(out, ref, "train B=256")
    n = 0
    for k, v in rloss.items():
        if k.endswith("batch_size"):
            continue
        got, want = float(loss[k]), float(v)
        assert abs(got - want) <= TOL * max(abs(want), 1e-2), (k, got, want)
        n += 1
    assert n >= 17
    worst, cf = _check_grads(model, sdg)
    print(f"B=256 train: worst per-tensor gradient cosine {worst[0]:.6f} ({worst[1]}), flat cosine {cf:.7f}, engines {used}")


def test_inference_b1024_benchmark_dispatch_vs_fp64_oracle():
    """BASELINE configs[1]: B=1024 eval forward (no-keep dual-sub-tile LSTM kernel, FP16 hand-off between the LSTM
    layers, BatchNorm running statistics, two streams), eager and through the CUDA-graph replay bench.py times."""
    B = 1024
    model, sd64 = _seq_model(seed=22)
    model.eval()
    batch = seq_inputs(B, TA, TV, TT, seed=22)
    dev = [cu(t) for t in batch[:5]]
    with torch.no_grad():
        out = model(*dev)
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.sequence_model(*batch[:5], sd64, training=False)
    _check_nig(out, ref, "infer B=1024")
    replay, gout = capture_forward(model, *dev)
    replay()
    torch.cuda.synchronize()
    _check_nig(gout, ref, "infer B=1024 (graph replay)")
    for k in ("mu_all", "nu", "alpha", "beta"):
        assert rel_l2(gout[k], out[k]) <= 1e-6, k


def test_sequence_trainer_graph_replay_matches_eager_and_oracle():
    """trainer.capture() of the SEQUENCE model (three-stream fork, deferred weight gradients, direct gradient
    accumulation, fused clip + AdamW) replays the same step the eager trainer runs; the captured step's loss and its
    flat gradient buffer are checked against the fp64 oracle too (B = 64: same kernels as B = 256, oracle in seconds)."""
    import copy
    B = 64
    base, sd64 = _seq_model(seed=23)
    base.train()
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    t1 = DEERDataParallelTrainer(m1, learning_rate=1e-4)
    t2 = DEERDataParallelTrainer(m2, learning_rate=1e-4)
    raw = seq_inputs(B, TA, TV, TT, seed=23)
    keys = ("audio_features", "video_features", "text_features", "attention_mask", "linguistic_features", "targets")
    batch = {k: cu(t) for k, t in zip(keys, raw)}
    # gradients of the first step, before the optimizer touches the weights
    l_eager = t1.forward_backward(batch).clone()
    g_eager = t1.flat.grads.clone()
    replay = t2.capture(batch, warmup=0)      # the capture pass launches nothing; weights still at their initial values
    l_graph = replay().clone()
    torch.cuda.synchronize()
    assert torch.allclose(l_eager, l_graph, rtol=1e-4, atol=1e-6), (l_eager, l_graph)
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd64.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    _, rloss = O.sequence_model_loss(*raw[:5], raw[5], sdg, training=True)
    rloss["total_loss"].backward()
    assert abs(float(l_graph[-1]) - float(rloss["total_loss"])) <= TOL * abs(float(rloss["total_loss"]))
    # flat gradient buffer of the eager trainer step vs the oracle (the replayed step already applied AdamW)
    names = t1.flat.names
    got, want = [], []
    for n, o in zip(names, t1.flat.offsets):
        og = sdg[n].grad
        if og is None:
            continue
        got.append(g_eager[o:o + og.numel()].double().cpu())
        want.append(og.flatten())
    assert cosine(torch.cat(got), torch.cat(want)) >= 0.999
    # a few more steps on both: the replayed trainer follows the eager one
    t1.optimizer_step()
    # (the first Adam steps move every weight by ~lr * sign(g): gradients whose sign is decided by the run-to-run noise of
    # the atomically accumulated / split-K reductions send the two trajectories apart by O(lr) per step, and the
    # bin-based ECE components react non-smoothly -- hence the loose tolerance on the later steps; step 1 above is tight)
    for i in range(3):
        a = t1.train_step(batch).clone()
        b = replay().clone()
        assert abs(float(a[-1]) - float(b[-1])) <= 2e-2 * abs(float(a[-1])), (i, a, b)
        assert torch.allclose(a, b, rtol=1e-1, atol=1e-3), (i, a, b)
    assert int(t1.step_tensor) == int(t2.step_tensor) == 4
    assert float((t1.flat.params - t2.flat.params).norm() / t1.flat.params.norm()) <= 2e-3


@pytest.mark.parametrize("B", [16384])
def test_pooled_model_large_batch_vs_fp64_oracle(B):
    """BASELINE configs[4]: the pooled CompleteDEERModel at a sweep batch size where every GEMM is large (M = B rows) --
    outputs, loss components and gradients against the fp64 oracle on whatever engine the sweep uses."""
    torch.manual_seed(0)
    cfg = deer_b200.ModelConfig(dropout=0.0)
    model = deer_b200.CompleteDEERModel(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd64 = det_state_dict(shapes, seed=31)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd64.items()})
    model = model.to(DEV).train()
    for mod in model.modules():      # UncertaintyEstimator has a hard-coded Dropout(0.2) (complete_project.py:193)
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    a, v, t, y = pooled_inputs(B, 31)
    out = model(cu(a), cu(v), cu(t))
    loss = model.compute_loss(out, cu(y))
    loss["total_loss"].backward()
    torch.cuda.synchronize()
    sdg = {k: (w.clone().requires_grad_(True) if w.is_floating_point() else w) for k, w in sd64.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref = O.pooled_model(a, v, t, sdg)
    rloss = O.multitask_deer_loss(O.pooled_loss_inputs(ref), y)
    rloss["total_loss"].backward()
    for d in DIMS:
        for k in ("mu", "nu", "alpha", "beta", "uncertainty"):
            assert_close(out[f"{d}_{k}"], ref[f"{d}_{k}"].detach(), TOL, f"{d}_{k}")
    assert_close(out["fused_features"], ref["fused_features"].detach(), TOL, "fused_features")
    for k, val in rloss.items():
        if k.endswith("batch_size"):
            continue
        got, want = float(loss[k]), float(val)
        assert abs(got - want) <= TOL * max(abs(want), 1e-2), (k, got, want)
    flat_c, flat_o = [], []
    for n, p in model.named_parameters():
        og = sdg[n].grad
        if og is None or float(og.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        if float(og.abs().max()) < 1e-12:
            continue
        c = cosine(p.grad, og)
        assert c >= 0.999, (n, c)
        flat_c.append(p.grad.flatten().double().cpu())
        flat_o.append(og.flatten())
    assert cosine(torch.cat(flat_c), torch.cat(flat_o)) >= 0.999


def test_sharded_loss_with_summed_statistics_equals_full_batch():
    """Exact global-batch loss semantics of the data-parallel trainer (trainer.py: `global_batch`, `stats_hook`): two
    shards whose phase-1 statistics are summed (what the all-reduce does) give the loss components and the evidence
    gradient of ONE full batch on one GPU, and both match the fp64 oracle."""
    from gen_common import nig_inputs
    B, W = 512, 2
    e64, y64 = nig_inputs(B, seed=41)
    e, y = cu(e64), cu(y64)
    full_l, full_g, _, _ = ops.nig_loss_raw(e, None, y, want_grad=True)
    # phase 1 on every shard first (a rank's hook sees the sum over ranks)
    n = B // W
    shard_stats = []
    for r in range(W):
        _, _, _, st = ops.nig_loss_raw(e[r * n:(r + 1) * n].contiguous(), None, y[r * n:(r + 1) * n].contiguous(),
                                       want_grad=False)
        shard_stats.append(st.clone())
    total = sum(shard_stats)

    def hook(stats):
        stats.copy_(total)

    grads, losses = [], []
    for r in range(W):
        l, g, _, _ = ops.nig_loss_raw(e[r * n:(r + 1) * n].contiguous(), None, y[r * n:(r + 1) * n].contiguous(),
                                      want_grad=True, stats_hook=hook, global_batch=B)
        grads.append(g)
        losses.append(l)
    torch.cuda.synchronize()
    assert_close(torch.cat(grads, 0), full_g, 1e-5, "sharded evidence gradient")
    for l in losses:
        assert torch.allclose(l, full_l, rtol=1e-5, atol=1e-7), (l, full_l)
    # and the oracle
    eg = e64.clone().requires_grad_(True)
    nig = O.nig_from_evidence(eg)
    pred = {f"{d}_{k}": nig[k][:, i:i + 1] for i, d in enumerate(DIMS) for k in ("mu", "nu", "alpha", "beta")}
    rl = O.multitask_deer_loss(pred, y64)
    rl["total_loss"].backward()
    assert abs(float(full_l[-1]) - float(rl["total_loss"])) <= TOL * abs(float(rl["total_loss"]))
    assert cosine(full_g, eg.grad) >= 0.99999


def _two_rank_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import deer_b200
    from deer_b200.trainer import DEERDataParallelTrainer, shard_batch
    from gen_common import det_state_dict, seq_inputs
    torch.manual_seed(0)
    model = deer_b200.SequenceDEERModel(dropout=0.0)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = det_state_dict(shapes, seed=51)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()})
    model = model.cuda().eval()    # eval(): BatchNorm on running statistics, so shards are exactly independent samples
    for p in model.parameters():
        p.requires_grad_(True)
    raw = seq_inputs(16, 24, 10, 12, seed=51)
    keys = ("audio_features", "video_features", "text_features", "attention_mask", "linguistic_features", "targets")
    full = {k: t.float().cuda() for k, t in zip(keys, raw)}
    tr = DEERDataParallelTrainer(model)
    losses = tr.forward_backward(shard_batch(full, rank, world))
    tr._allreduce(tr.flat.grads)
    torch.cuda.synchronize()
    q.put((rank, losses.cpu().numpy(), tr.flat.grads.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_step_equals_single_rank_full_batch():
    """N-rank == 1-rank: two NCCL ranks on half batches (exact global loss statistics + summed gradients) reproduce the
    loss and the flat gradient buffer of one rank on the full batch."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    # single rank, full batch
    torch.manual_seed(0)
    model = deer_b200.SequenceDEERModel(dropout=0.0)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = det_state_dict(shapes, seed=51)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()})
    model = model.cuda().eval()
    raw = seq_inputs(16, 24, 10, 12, seed=51)
    keys = ("audio_features", "video_features", "text_features", "attention_mask", "linguistic_features", "targets")
    full = {k: t.float().cuda() for k, t in zip(keys, raw)}
    tr = DEERDataParallelTrainer(model)
    l1 = tr.forward_backward(full).cpu().numpy()
    g1 = tr.flat.grads.cpu().numpy()
    for rank, l, g in res:
        np.testing.assert_allclose(l, l1, rtol=2e-4, atol=1e-6)
        c = float(np.dot(g.astype(np.float64), g1.astype(np.float64)) /
                  (np.linalg.norm(g.astype(np.float64)) * np.linalg.norm(g1.astype(np.float64))))
        assert c >= 0.99999, (rank, c)
        assert abs(np.linalg.norm(g) / np.linalg.norm(g1) - 1.0) <= 1e-3
