"""One warm training step (+ optional inference step) of the sequence DEER model, for ncu launch lists.
usage: profile_step.py [train|infer] [B]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import deer_b200
from deer_b200.trainer import DEERDataParallelTrainer
sys.path.insert(0, ROOT)
from bench import synth_batch

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = deer_b200.SequenceDEERModel(dropout=0.3).to(dev)
gen = torch.Generator().manual_seed(1)
batch = synth_batch(B, dev, gen)
if mode == "train":
    model.train()
    tr = DEERDataParallelTrainer(model)
    for _ in range(steps):
        tr.train_step(batch)
else:
    model.eval()
    with torch.no_grad():
        for _ in range(steps):
            model(batch["audio_features"], batch["video_features"], batch["text_features"], batch["attention_mask"],
                  batch["linguistic_features"])
torch.cuda.synchronize()
print("done", mode, B)
