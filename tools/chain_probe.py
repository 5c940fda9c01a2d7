"""GPU probe: fusion + head on the fused 3xTF32 nodes vs the SIMT engine, per tensor, at several batch sizes with
deterministic non-trivial weights (tests/golden/gen_common.py)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import deer_b200
from deer_b200 import ops
from gen_common import det_state_dict

DEV = "cuda"
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))

for B in (4, 7, 32, 33, 96):
    g = torch.Generator().manual_seed(B)
    a, v, t = (torch.randn(B, 512, generator=g).to(DEV) for _ in range(3))
    y = torch.tanh(torch.randn(B, 3, generator=g)).to(DEV)
    res = {}
    for eng in (ops.ENGINE_SIMT, ops.ENGINE_X3):
        ops.set_exact_engine(eng)
        torch.manual_seed(0)
        fus = deer_b200.HierarchicalMultimodalFusion(512, 512, 512, fusion_dim=512, intermediate_dim=256, dropout=0.0)
        head = deer_b200.MultiDimensionalDEER(512, 3, 256, 0.0)
        for m in (fus, head):
            sd = det_state_dict({k: tuple(x.shape) for k, x in m.state_dict().items()}, seed=3)
            m.load_state_dict({k: x.float() for k, x in sd.items()})
            m.to(DEV).train()
        ins = [x.clone().requires_grad_(True) for x in (a, v, t)]
        ops.begin_step()
        f = fus(*ins)
        inter = {"av": f["audiovisual_features"], "tri": f["trimodal_features"], "fused": f["fused_features"]}
        for x in inter.values():
            x.retain_grad()
        out = head(f["fused_features"])
        loss = deer_b200.MultiTaskDEERLoss()(out, y)
        loss["total_loss"].backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.clone() for n, p in list(fus.named_parameters()) + list(head.named_parameters()) if p.grad is not None}
        res[eng] = (grads, [x.grad.clone() for x in ins], {k: x.grad.clone() for k, x in inter.items()},
                    {k: x.detach().clone() for k, x in inter.items()})
    s, x = res[ops.ENGINE_SIMT], res[ops.ENGINE_X3]
    print(f"B={B}: fwd " + " ".join(f"{k}={rel(x[3][k], s[3][k]):.1e}" for k in s[3]) +
          " | dgrad " + " ".join(f"{k}={rel(x[2][k], s[2][k]):.1e}" for k in s[2]) +
          " | din " + " ".join(f"{rel(p, q):.1e}" for p, q in zip(x[1], s[1])))
    bad = sorted(((rel(x[0][n], s[0][n]), n) for n in s[0] if float(s[0][n].abs().max()) > 0), reverse=True)[:6]
    print("   worst params: " + " ".join(f"{n}={r:.1e}" for r, n in bad), flush=True)
