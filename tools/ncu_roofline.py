"""The roofline kernels of bench.py, a few launches each (for `ncu --set full`): the layer-1 LSTM input projection with
its FP16 output on the CTA-pair 16-bit GEMM, and the two fused NIG head + loss passes at B = 2^22."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deer_b200  # noqa
from deer_b200 import ops
M, N, K = 76800, 2048, 512
nbuf = 3
A = [(torch.randn(M, K, device="cuda") * 0.5).half() for _ in range(nbuf)]
W = (torch.randn(N, K, device="cuda") * 0.05).half()
bias = torch.randn(N, device="cuda")
C16 = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(nbuf)]
for i in range(4):
    ops.gemm_h16(A[i % nbuf], K, 0, W, K, 1, None, 0, M, N, K, bias=bias, C16=C16[i % nbuf], ldc16=N)
torch.cuda.synchronize()
del A, C16
n = 1 << 22
ev = [torch.randn(n, 3, 4, device="cuda") for _ in range(2)]
tg = [torch.tanh(torch.randn(n, 3, device="cuda")) for _ in range(2)]
for i in range(3):
    ops.nig_loss_raw(ev[i % 2], None, tg[i % 2], want_nig=True, want_grad=True)
torch.cuda.synchronize()
print("ok")
