"""GPU probe: does a tcgen05 GEMM CTA that fits beside a recurrence CTA (gemm_tf32_kernel: 128x128 tiles, 100 KB of shared
memory, 128 TMEM columns) actually run there, and what does it cost the recurrence?  Times the BPTT kernel (B=256, T=300)
alone, a train of TF32 GEMMs alone, and both at once on two streams."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200  # noqa: E402,F401
from deer_b200 import _lib, ops  # noqa: E402
from deer_b200._lib import call, ptr  # noqa: E402

dev = "cuda"
T, B, H = 300, 256, 256
Bp = (B + 31) // 32 * 32
w = [torch.randn(4 * H, H, device=dev) * 0.05 for _ in range(2)]
gact = (torch.rand(T * 2 * Bp * 4 * H, device=dev) * 0.9).half()
c = (torch.randn(T * 2 * Bp * H, device=dev) * 0.5).half()
dh = torch.randn(T, B, 2 * H, device=dev) * 1e-3
dpre16 = torch.empty(T, B, 2, 4 * H, device=dev, dtype=torch.bfloat16)
db = torch.zeros(2, 4 * H, device=dev)


def lstm():
    call("deer_lstm_cluster_bwd", ptr(gact), ptr(c), ptr(dh), ptr(w[0]), ptr(w[1]), None, ptr(db), dpre16.data_ptr(), T, B, H)


M, N, K = 12800, 512, 1536
A = torch.randn(M, K, device=dev)
Wt = torch.randn(N, K, device=dev) * 0.05
C = torch.empty(M, N, device=dev)


def gemms(n):
    for _ in range(n):
        ops.gemm(A, K, 0, Wt, K, 1, C, N, M, N, K)


def timed(fn, stream, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


sa, sb = torch.cuda.Stream(priority=-1), torch.cuda.Stream()
for pair in (1, 0):
    _lib.set_option(6, pair)          # DEER_OPT_TF32_PAIR: 1 = CTA-pair 256x256 kernel (225 KB), 0 = 128x128 kernel (100 KB)
    t_l = timed(lstm, sa)
    n = 12
    t_g = timed(lambda: gemms(n), sb)
    # both: GEMM train on sb, LSTM on sa, started together
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(sb):
        e[2].record()
        gemms(n)
        e[3].record()
    with torch.cuda.stream(sa):
        e[0].record()
        lstm()
        e[1].record()
    torch.cuda.synchronize()
    both_l, both_g = e[0].elapsed_time(e[1]) * 1e3, e[2].elapsed_time(e[3]) * 1e3
    print(f"tf32_pair={pair}: BPTT alone {t_l:7.1f} us | {n} GEMMs alone {t_g:7.1f} us ({t_g / n:5.1f} each) | together: BPTT "
          f"{both_l:7.1f} us, GEMM train {both_g:7.1f} us  (serial sum {t_l + t_g:7.1f}, wall {max(both_l, both_g):7.1f})", flush=True)
