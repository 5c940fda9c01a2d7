"""GPU probe: the 3xTF32 engine against the SIMT engine on the scorer / chain shapes of a small batch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200
from deer_b200 import ops

DEV = "cuda"
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
g = torch.Generator().manual_seed(0)
for (M, D, Hd) in ((1200, 512, 256), (200, 512, 256), (256, 768, 384), (4, 512, 512), (48, 512, 256)):
    x = torch.randn(M, D, generator=g).to(DEV)
    w1 = (torch.randn(Hd, D, generator=g) * 0.05).to(DEV)
    b1 = torch.randn(Hd, generator=g).to(DEV)
    dh = torch.randn(M, Hd, generator=g).to(DEV)
    res = {}
    for eng in (ops.ENGINE_SIMT, ops.ENGINE_X3):
        hidden = torch.empty(M, Hd, device=DEV)
        ops.gemm(x, D, 0, w1, D, 1, hidden, Hd, M, Hd, D, bias=b1, act=2, engine=eng)
        dx = torch.ones(M, D, device=DEV)
        ops.gemm(dh, Hd, 0, w1, D, 0, dx, D, M, D, Hd, beta=1.0, engine=eng)
        dw = torch.ones(Hd, D, device=DEV)
        ops.gemm(dh, Hd, 1, x, D, 0, dw, D, Hd, D, M, beta=1.0, engine=eng)
        torch.cuda.synchronize()
        res[eng] = (hidden, dx, dw)
    r64 = (torch.tanh(x.double() @ w1.double().t() + b1.double()), 1 + dh.double() @ w1.double(), 1 + dh.double().t() @ x.double())
    print(f"M={M} D={D} Hd={Hd}: " + " ".join(
        f"{n}: x3-vs-simt {rel(res[ops.ENGINE_X3][i], res[ops.ENGINE_SIMT][i]):.1e} x3-vs-f64 {rel(res[ops.ENGINE_X3][i], r64[i]):.1e} simt-vs-f64 {rel(res[ops.ENGINE_SIMT][i], r64[i]):.1e}"
        for i, n in enumerate(("hidden", "dx", "dw"))), flush=True)
