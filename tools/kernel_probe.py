"""Launch each dominant kernel once or twice at its bench size (for `ncu --set full`): the fused NIG head+loss pair at
2^22 samples, the layer-1 input-projection GEMM, one persistent LSTM forward and one BPTT (B=256, T=300)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200  # noqa: E402,F401
from deer_b200 import ops  # noqa: E402
from deer_b200._lib import call, ptr  # noqa: E402

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = 2
if which in ("all", "nig"):
    n = 1 << 22
    ev = torch.randn(n, 3, 4, device=dev)
    tg = torch.tanh(torch.randn(n, 3, device=dev))
    for _ in range(reps):
        ops.nig_loss_raw(ev, None, tg, want_nig=True, want_grad=True)
    del ev, tg
if which in ("all", "gemm"):
    M, N, K = 76800, 1024, 512
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    C = torch.empty(M, 2 * N, device=dev)
    for _ in range(reps):
        ops.gemm(A, K, 0, W, K, 1, C, 2 * N, M, N, K, bias=b)
    del A, C
if which in ("all", "lstm"):
    T, B, H = 300, 256, 256
    Bp = (B + 31) // 32 * 32
    pre = torch.randn(T, B, 2, 4 * H, device=dev)
    w = [torch.randn(4 * H, H, device=dev) * 0.05 for _ in range(2)]
    h = torch.empty(T, B, 2 * H, device=dev)
    gact = torch.empty(T * 2 * Bp * 4 * H, device=dev)
    c = torch.empty(T * 2 * Bp * H, device=dev)
    dh = torch.randn(T, B, 2 * H, device=dev) * 1e-3
    dpre = torch.empty(T, B, 2, 4 * H, device=dev)
    db = torch.zeros(2, 4 * H, device=dev)
    for _ in range(reps):
        call("deer_lstm_cluster_fwd", ptr(pre), ptr(w[0]), ptr(w[1]), ptr(h), ptr(gact), ptr(c), None, None, T, B, H)
        call("deer_lstm_cluster_bwd", ptr(gact), ptr(c), ptr(dh), ptr(w[0]), ptr(w[1]), ptr(dpre), ptr(db), None, T, B, H)
torch.cuda.synchronize()
print("kernel_probe done", which)
