"""Error of the NIG outputs / loss / gradients of the sequence composite vs the fp64 oracle under each engine policy."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import deer_b200
from deer_b200 import ops
from gen_common import det_state_dict, seq_inputs
from oracle import deer_oracle as O
from helpers import rel_l2, cosine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = deer_b200.SequenceDEERModel(dropout=0.0)
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
sd64 = det_state_dict(shapes, seed=21)
model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd64.items()})
model = model.cuda().train()
batch = seq_inputs(B, 300, 50, 64, seed=21)
dev = [t.float().cuda() for t in batch]
sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd64.items()}
ref, rloss = O.sequence_model_loss(*batch[:5], batch[5], sdg, training=True)
ref["gamma"] = ref["mu_all"]
for k in ("nu", "alpha", "beta"):
    ref[k] = torch.cat([ref[f"{d}_{k}"] for d in O.DIMS], dim=1)
rloss["total_loss"].backward()

def run(tag):
    model.zero_grad(set_to_none=True)
    out = model(*dev[:5])
    loss = model.compute_loss(out, dev[5])
    loss["total_loss"].backward()
    errs = {k: rel_l2(out[k], ref[k]) for k in ("gamma", "nu", "alpha", "beta", "uncertainty_all", "fused_features",
                                                 "audio_encoded", "video_encoded", "text_encoded")}
    lerr = abs(float(loss["total_loss"]) - float(rloss["total_loss"])) / abs(float(rloss["total_loss"]))
    cs = []
    for n, p in model.named_parameters():
        og = sdg[n].grad
        if og is None or float(og.abs().max()) < 1e-12:
            continue
        cs.append((cosine(p.grad, og), n))
    print(tag, " ".join(f"{k}={v:.1e}" for k, v in errs.items()), f"loss={lerr:.1e}", f"min grad cos={min(cs)[0]:.6f} ({min(cs)[1]})")

ops.set_gemm_engine(ops.ENGINE_SIMT); run("simt-fp32       :")
ops.set_gemm_engine(ops.ENGINE_AUTO); ops.set_exact_small_forward(False); run("tf32 everywhere :")
ops.set_exact_small_forward(True, backward=False); run("tf32 big + fp32 small fwd:")
ops.set_exact_small_forward(True); run("tf32 big + fp32 small fwd+bwd:")
from deer_b200 import _lib
_lib.set_option(1, 0); run("same, TMA raw fp32 (truncate):"); _lib.set_option(1, 1)
