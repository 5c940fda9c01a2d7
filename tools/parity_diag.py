"""GPU diagnostic: per-group gradient cosines of a training step against the fp64 oracle under several engine policies."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import deer_b200
from deer_b200 import ops, _lib
from gen_common import det_state_dict, seq_inputs
from oracle import deer_oracle as O

DEV = "cuda"
def cos(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))
def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2].split(",") if len(sys.argv) > 2 else None
torch.manual_seed(0)
model = deer_b200.SequenceDEERModel(dropout=0.0)
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
sd64 = det_state_dict(shapes, seed=21)
model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd64.items()})
model = model.to(DEV).train()
batch = seq_inputs(B, 300, 50, 64, seed=21)
dev = [t.float().to(DEV) for t in batch]
sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd64.items()}
torch.set_num_threads(os.cpu_count())
ref, rloss = O.sequence_model_loss(*batch[:5], batch[5], sdg, training=True)
for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features"):
    ref[k].retain_grad()
rloss["total_loss"].backward()

def reset():
    ops.set_gemm_engine(ops.ENGINE_AUTO); ops.set_exact_engine(ops.ENGINE_X3)
    ops.set_exact_small_forward(True, 8192); ops.set_conv_window(True); ops.set_conv_exact(False, False)
    ops.set_lstm_gemm16(True); ops.set_lstm_pre16(True)

def run(tag, setup):
    reset()
    setup()
    model.zero_grad(set_to_none=True)
    out = model(*dev[:5])
    for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features"):
        out[k].retain_grad()
    loss = model.compute_loss(out, dev[5])
    loss["total_loss"].backward()
    torch.cuda.synchronize()
    groups = {}
    for n, p in model.named_parameters():
        og = sdg[n].grad
        if og is None or float(og.abs().max()) < 1e-12:
            continue
        g = n.split(".")[0]
        if g == "audio_encoder" and ".lstm." in n:
            g = "audio_lstm"
        c = cos(p.grad, og)
        w = groups.setdefault(g, [1.0, None, 0.0])
        if c < w[0]:
            w[0], w[1] = c, n
        w[2] = max(w[2], rel(p.grad, og))
    fc = cos(torch.cat([p.grad.flatten() for n, p in model.named_parameters() if sdg[n].grad is not None]),
             torch.cat([sdg[n].grad.flatten() for n, p in model.named_parameters() if sdg[n].grad is not None]))
    print(f"== {tag}: loss rel {abs(float(loss['total_loss'].detach()) - float(rloss['total_loss'].detach())) / abs(float(rloss['total_loss'].detach())):.2e} "
          f"mu rel {rel(out['mu_all'], ref['mu_all']):.2e} flat cos {fc:.6f}")
    print("   fwd rel: " + " ".join(f"{k}={rel(out[k], ref[k]):.1e}" for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features")))
    print("   upstream grad cos: " + " ".join(f"{k}={cos(out[k].grad, ref[k].grad):.6f}" for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features")))
    for g, (c, n, r) in sorted(groups.items()):
        print(f"   {g:14s} worst cos {c:.6f} (max rel {r:.1e}) {n}")
    sys.stdout.flush()

variants = {
    "default": lambda: None,
    "exact_scorers": lambda: ops.set_exact_small_forward(True, 10**9),
    "exact_scorers_fwd_only": lambda: ops.set_exact_small_forward(True, 10**9, small_rows_bwd=8192),
    "exact_scorers_bwd_only": lambda: ops.set_exact_small_forward(True, 8192, small_rows_bwd=10**9),
    "exact_scorers+conv": lambda: (ops.set_exact_small_forward(True, 10**9), ops.set_conv_exact(True, True)),
    "exact_conv_only": lambda: ops.set_conv_exact(True, True),
    "exact_all+lstm_tf32gemm": lambda: (ops.set_exact_small_forward(True, 10**9), ops.set_conv_exact(True, True), ops.set_lstm_gemm16(False)),
    "exact_engine_simt": lambda: ops.set_exact_engine(ops.ENGINE_SIMT),
    "all_simt": lambda: ops.set_gemm_engine(ops.ENGINE_SIMT),
}
for k, f in variants.items():
    if which is None and k == "all_simt" and B > 64:
        continue
    if which is None or k in which:
        run(k, f)
