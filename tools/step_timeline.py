"""Timeline of one CUDA-graph-replayed training step (and inference forward): timestamp kernels on every stream at the
encoder boundaries (forward and, through autograd, backward), printed in microseconds from the start of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deer_b200
from deer_b200 import ops
from deer_b200.trainer import DEERDataParallelTrainer, capture_forward
from bench import synth_batch

# under torchrun (WORLD_SIZE > 1): one rank per GPU, NCCL; every rank prints its own timeline (marks around the loss
# statistics all-reduce = head_out .. loss_done, and the gradient exchange = backward_done .. grads_exchanged)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = deer_b200.SequenceDEERModel(dropout=0.3).to(dev).train()
tr = DEERDataParallelTrainer(model)
if "--overlap-exchange" in sys.argv:
    tr.overlap_exchange = True
gen = torch.Generator().manual_seed(1 + rank)
batch = synth_batch(256, dev, gen)
if "--input-grads" in sys.argv:   # marks at the END of each encoder's backward (costs the three input-gradient GEMMs)
    for k in ("audio_features", "video_features", "text_features"):
        batch[k].requires_grad_(True)
for _ in range(2):
    tr.train_step(batch)
ops.timeline_begin(dev)
_orig = tr.optimizer_step
def opt():
    ops.mark("backward_done")
    _orig()
    ops.mark("optimizer_done")
tr.optimizer_step = opt
replay = tr.capture(batch, warmup=1)
starts = []
for _ in range(6):
    replay()
    torch.cuda.synchronize()
    starts.append(dict(ops.timeline_read())["step_start"])
for _ in range(4):      # back-to-back replays: the steady-state period of a step
    replay()
torch.cuda.synchronize()
marks = ops.timeline_read()
t0 = dict(marks)["step_start"]
lines = [f"rank {rank}/{world}: training step B=256 (graph replay), us from step start:"]
for n, v in sorted(marks, key=lambda kv: kv[1]):
    lines.append(f"  {(v - t0) / 1e3:9.1f}  {n}")
print("\n".join(lines), flush=True)
ops.timeline_end()
if world > 1:
    from bench import teardown_distributed     # graph-captured collectives: destroy_process_group under a watchdog
    teardown_distributed(torch, dist, tr)
