"""Timeline of one CUDA-graph-replayed training step (and inference forward): timestamp kernels on every stream at the
encoder boundaries (forward and, through autograd, backward), printed in microseconds from the start of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deer_b200
from deer_b200 import ops
from deer_b200.trainer import DEERDataParallelTrainer, capture_forward
from bench import synth_batch

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = deer_b200.SequenceDEERModel(dropout=0.3).to(dev).train()
tr = DEERDataParallelTrainer(model)
gen = torch.Generator().manual_seed(1)
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
for _ in range(3):
    replay()
torch.cuda.synchronize()
marks = ops.timeline_read()
t0 = dict(marks)["step_start"]
print("training step B=256 (graph replay), us from step start:")
for n, v in sorted(marks, key=lambda kv: kv[1]):
    print(f"  {(v - t0) / 1e3:9.1f}  {n}")
ops.timeline_end()
