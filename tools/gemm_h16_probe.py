"""Timing of the 16-bit tcgen05 engine on the LSTM layer-1 GEMM shapes vs the TF32 engine (CUDA events)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200  # noqa
from deer_b200 import ops

dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

MT, H = 76800, 256
cases = [
    ("fwd proj  [MT,512]x[1024,512]^T", MT, 1024, 512, 0, 1),
    ("fwd proj0 [MT,88]x[1024,88]^T  ", MT, 1024, 88, 0, 1),
    ("dx        [MT,1024]x[1024,512] ", MT, 512, 1024, 0, 0),
    ("dW_ih     [MT,1024]^T x [MT,512]", 1024, 512, MT, 1, 0),
    ("dW_hh     [MT,1024]^T x [MT,256]", 1024, 256, MT, 1, 0),
]
from deer_b200 import _lib
lib = _lib.load()
for name, M, N, K, ta, tb in cases:
    A32 = torch.randn((K, M) if ta else (M, K), device=dev) * 0.1
    B32 = torch.randn((N, K) if tb else (K, N), device=dev) * 0.1
    A16, B16 = A32.half(), B32.half()
    C = torch.zeros(M, N, device=dev)
    beta = 1.0 if ta else 0.0
    t16 = timeit(lambda: ops.gemm_h16(A16, A16.shape[1], ta, B16, B16.shape[1], tb, C, N, M, N, K, beta=beta))
    t32 = timeit(lambda: ops.gemm(A32, A32.shape[1], ta, B32, B32.shape[1], tb, C, N, M, N, K, beta=beta, engine=ops.ENGINE_TF32))
    buf = torch.zeros(8, dtype=torch.int64, device=dev)
    lib.deer_gemm_h16_set_profile_buffer(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm_h16(A16, A16.shape[1], ta, B16, B16.shape[1], tb, C, N, M, N, K, beta=beta)
    e1.record(); torch.cuda.synchronize()
    lib.deer_gemm_h16_set_profile_buffer(None)
    v = buf.cpu().tolist()
    print(f"   block0 cycles: producer-wait-empty {v[0]} | mma-wait-tmem_empty {v[1]} | mma-wait-full {v[2]} | epi-wait-tmem_full {v[3]} | epi-busy {v[4]} (tmem_ld {v[5]}, sts {v[6]}, readback+stg {v[7]}) | kernel {e0.elapsed_time(e1)*1e3:.0f} us")
    fl = 2.0 * M * N * K
    print(f"{name}: h16 {t16:8.1f} us = {fl / t16 / 1e6:7.1f} TFLOP/s | tf32 {t32:8.1f} us = {fl / t32 / 1e6:7.1f} TFLOP/s", flush=True)
