"""2-GPU diagnostic: reduced gradients with the early (overlapped) exchange vs one all-reduce at the end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import deer_b200
from deer_b200.trainer import DEERDataParallelTrainer
from bench import synth_batch
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
model = deer_b200.SequenceDEERModel(dropout=0.3).to(dev).train()
tr = DEERDataParallelTrainer(model)
b = synth_batch(256, dev, torch.Generator().manual_seed(1234 + rank))
def run(ov):
    tr.overlap_exchange = ov
    tr.forward_backward(b)
    if not tr._grads_reduced:
        tr._allreduce(tr.flat.grads)
    tr._grads_reduced = False
    torch.cuda.synchronize()
    return tr.flat.grads.clone()
run(False)
g = {k: run(v) for k, v in (("f1", False), ("f2", False), ("t1", True), ("t2", True))}
def rel(a, c): return float((a - c).norm() / c.norm())
if rank == 0:
    print("f1-f2", rel(g["f1"], g["f2"]), "t1-t2", rel(g["t1"], g["t2"]), "t1-f1", rel(g["t1"], g["f1"]), "audio_end", tr._audio_end, flush=True)
    worst = []
    offs = tr.flat.offsets + [tr.flat.numel]
    for i, n in enumerate(tr.flat.names):
        a, c = g["t1"][offs[i]:offs[i + 1]], g["f1"][offs[i]:offs[i + 1]]
        d0 = g["f2"][offs[i]:offs[i + 1]]
        if float(c.norm()) > 0:
            worst.append((rel(a, c), rel(d0, c), n))
    worst.sort(reverse=True)
    for w in worst[:8]:
        print("  overlap-vs-end %.2e   end-vs-end %.2e   %s" % w, flush=True)
dist.destroy_process_group()
