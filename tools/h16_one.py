"""One shape of the 16-bit GEMM, a few launches (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deer_b200  # noqa
from deer_b200 import ops
M, N, K = 76800, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 512
A = (torch.randn(M, K, device="cuda") * 0.1).half()
B = (torch.randn(N, K, device="cuda") * 0.1).half()
bias = torch.randn(N, device="cuda")
C = torch.zeros(M, 2 * N, device="cuda")
for _ in range(4):
    ops.gemm_h16(A, K, 0, B, K, 1, C, 2 * N, M, N, K, bias=bias)
torch.cuda.synchronize()
print("ok")
