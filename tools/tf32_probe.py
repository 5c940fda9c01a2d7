"""TF32 GEMM engines on the model's time-batched shapes: 128x128-tile kernel vs the CTA-pair kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deer_b200  # noqa
from deer_b200 import ops, _lib

SHAPES = [  # name, M, N, K, ta, tb, act, beta
    ("audio scorer fwd", 76800, 256, 512, 0, 1, 2, 0.0), ("audio scorer dX", 76800, 512, 256, 0, 0, 0, 0.0),
    ("audio scorer dW", 256, 512, 76800, 1, 0, 0, 1.0),
    ("video spatial fwd", 12800, 512, 256, 0, 1, 1, 0.0), ("video conv fwd", 12800, 512, 1536, 0, 1, 0, 0.0),
    ("video conv dX", 12800, 1536, 512, 0, 0, 0, 0.0), ("video conv dW", 512, 1536, 12800, 1, 0, 0, 1.0),
    ("video scorer fwd", 12800, 256, 512, 0, 1, 2, 0.0),
    ("text scorer fwd", 16384, 384, 768, 0, 1, 2, 0.0), ("text scorer dX", 16384, 768, 384, 0, 0, 0, 0.0),
    ("text scorer dW", 384, 768, 16384, 1, 0, 0, 1.0),
]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


tot = [0.0, 0.0]
for name, M, N, K, ta, tb, act, beta in SHAPES:
    A = torch.randn((K, M) if ta else (M, K), device="cuda") * 0.1
    B = torch.randn((N, K) if tb else (K, N), device="cuda") * 0.1
    C = torch.zeros(M, N, device="cuda")
    bias = torch.randn(N, device="cuda") if beta == 0 else None
    t = []
    for pair in (0, 1):
        _lib.set_option(6, pair)
        t.append(timeit(lambda: ops.gemm(A, A.shape[1], ta, B, B.shape[1], tb, C, N, M, N, K, bias=bias, act=act,
                                         beta=beta, engine=ops.ENGINE_TF32)))
        tot[pair] += t[-1]
    fl = 2.0 * M * N * K
    print(f"{name:18s} [{M},{N},{K}] ta={ta} tb={tb}: 128x128 {t[0]:7.1f} us ({fl/t[0]/1e6:6.1f} TF)   pair {t[1]:7.1f} us "
          f"({fl/t[1]/1e6:6.1f} TF)", flush=True)
print(f"total: 128x128 {tot[0]:.0f} us, pair {tot[1]:.0f} us")
