"""GPU probe: the whole sequence model at a small batch, fused 3xTF32 nodes vs SIMT engine, gradients at the fusion
boundaries and per-parameter, several seeds."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import deer_b200
from deer_b200 import ops
from gen_common import det_state_dict, seq_inputs

DEV = "cuda"
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for seed in (21, 5, 7):
    torch.manual_seed(0)
    model = deer_b200.SequenceDEERModel(dropout=0.0)
    sd64 = det_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd64.items()})
    model = model.to(DEV).train()
    batch = [t.float().to(DEV) for t in seq_inputs(B, 300, 50, 64, seed=seed)]
    stash = {}
    orig = model.fusion.forward
    def wrapped(a, v, t, uncertainties=None):
        r = orig(a, v, t, uncertainties)
        stash.update(a=a, v=v, t=t, av=r["audiovisual_features"], tri=r["trimodal_features"], fused=r["fused_features"])
        for x in stash.values():
            x.retain_grad()
        return r
    model.fusion.forward = wrapped
    res = {}
    for streams in (True, False):
        for eng in (ops.ENGINE_SIMT, ops.ENGINE_X3):
            ops.set_exact_engine(eng)
            ops.set_branch_streams(streams)
            model.zero_grad(set_to_none=True)
            out = model(*batch[:5])
            loss = model.compute_loss(out, batch[5])
            loss["total_loss"].backward()
            torch.cuda.synchronize()
            res[(streams, eng)] = ({k: x.grad.clone() for k, x in stash.items()}, {k: x.detach().clone() for k, x in stash.items()},
                                   {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    ref = res[(False, ops.ENGINE_SIMT)]
    for key, r in res.items():
        if key == (False, ops.ENGINE_SIMT):
            continue
        worst = sorted(((rel(r[2][n], ref[2][n]), n) for n in ref[2] if float(ref[2][n].abs().max()) > 0), reverse=True)[:3]
        print(f"seed {seed} streams={key[0]} engine={'x3' if key[1] == ops.ENGINE_X3 else 'simt'}: fwd " +
              " ".join(f"{k}={rel(r[1][k], ref[1][k]):.1e}" for k in ref[1]) + " | grad " +
              " ".join(f"{k}={rel(r[0][k], ref[0][k]):.1e}" for k in ref[0]) + " | worst " +
              " ".join(f"{n}={e:.1e}" for e, n in worst), flush=True)
