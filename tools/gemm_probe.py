"""Numerics + speed probe of the tcgen05 GEMM engine against fp64 and the fp32 SIMT engine (run on a B200)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deer_b200
from deer_b200 import ops, _lib

dev = "cuda"
torch.manual_seed(0)


def run(M, N, K, ta, tb, engine, act=0, beta=0.0, bias=True):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g) * 0.1
    bv = torch.randn(N, generator=g) if bias else None
    C0 = torch.randn(M, N, generator=g)
    C = C0.to(dev).clone()
    ops.gemm(A.to(dev), A.shape[1], ta, B.to(dev), B.shape[1], tb, C, N, M, N, K, bias=None if bv is None else bv.to(dev),
             act=act, beta=beta, engine=engine)
    torch.cuda.synchronize()
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
    if bv is not None:
        ref = ref + bv.double()
    ref = ref + beta * C0.double()
    if act == 1:
        ref = torch.relu(ref)
    if act == 2:
        ref = torch.tanh(ref)
    d = C.double().cpu() - ref
    return float(d.norm() / ref.norm()), float(d.mean() / ref.abs().mean())


shapes = [(128, 128, 32, 0, 1), (128, 128, 64, 0, 1), (256, 256, 128, 0, 1), (300, 200, 84, 0, 1), (256, 512, 512, 0, 1),
          (128, 128, 64, 0, 0), (256, 256, 128, 0, 0), (300, 200, 100, 0, 0),
          (128, 128, 64, 1, 0), (256, 384, 4096, 1, 0), (1024, 256, 76800 // 4, 1, 0),
          (128, 128, 64, 1, 1), (200, 136, 96, 1, 1)]
for rnd in (1, 0):
    _lib.set_option(1, rnd)
    print(f"--- TMA tf32 rounding = {rnd}")
    for (M, N, K, ta, tb) in shapes:
        try:
            e, bias_ = run(M, N, K, ta, tb, ops.ENGINE_TF32)
            e2, _ = run(M, N, K, ta, tb, ops.ENGINE_TF32, act=2, beta=1.0)
            print(f"M={M:5d} N={N:5d} K={K:6d} ta={ta} tb={tb}: rel={e:.2e} meanbias={bias_:+.2e} (tanh,beta1 rel={e2:.2e})")
        except Exception as ex:
            print(f"M={M} N={N} K={K} ta={ta} tb={tb}: FAILED {ex}")
            raise
_lib.set_option(1, 1)

# speed: the big time-batched shapes
def bench(M, N, K, ta, tb, engine, beta=0.0, n=20):
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    for _ in range(3):
        ops.gemm(A, A.shape[1], ta, B, B.shape[1], tb, C, N, M, N, K, beta=beta, engine=engine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.gemm(A, A.shape[1], ta, B, B.shape[1], tb, C, N, M, N, K, beta=beta, engine=engine)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, 2.0 * M * N * K / ms / 1e9


for name, (M, N, K, ta, tb, beta) in {
        "lstm_l1_inproj fwd [76800,512]x[2048,512]^T": (76800, 2048, 512, 0, 1, 0.0),
        "lstm_l0_inproj fwd [76800,84]x[2048,84]^T": (76800, 2048, 84, 0, 1, 0.0),
        "scorer fwd [76800,512]x[256,512]^T": (76800, 256, 512, 0, 1, 0.0),
        "lstm_l1 dgrad [76800,2048]x[2048,512]": (76800, 512, 2048, 0, 0, 0.0),
        "lstm_l1 wgrad [2048,76800]x[76800,512]": (2048, 512, 76800, 1, 0, 1.0),
        "whh wgrad [1024,76544]x[76544,256]": (1024, 256, 76544, 1, 0, 1.0),
        "conv [12800,1536]x[512,1536]^T": (12800, 512, 1536, 0, 1, 0.0),
        "small [256,512]x[512,512]^T": (256, 512, 512, 0, 1, 0.0)}.items():
    ms, tf = bench(M, N, K, ta, tb, ops.ENGINE_TF32, beta)
    ms2, tf2 = bench(M, N, K, ta, tb, ops.ENGINE_SIMT, beta, n=3)
    print(f"{name:48s} tcgen05 {ms:8.3f} ms {tf:8.1f} TFLOP/s | simt {ms2:8.3f} ms {tf2:6.1f} TFLOP/s")
