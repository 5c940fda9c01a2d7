import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deer_b200 import ops
dev = "cuda"
for (M, N, K, tb, beta) in [(256, 1024, 256, 1, 1.0), (1024, 1024, 256, 1, 1.0), (256, 256, 1024, 0, 0.0), (256, 512, 512, 1, 0.0)]:
    A = torch.randn(M, K, device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.zeros(M, N, device=dev)
    for eng in (ops.ENGINE_TF32, ops.ENGINE_SIMT):
        for _ in range(5):
            ops.gemm(A, K, 0, B, B.shape[1], tb, C, N, M, N, K, beta=beta, engine=eng)
        torch.cuda.synchronize()
        n = 300
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(n):
            ops.gemm(A, K, 0, B, B.shape[1], tb, C, N, M, N, K, beta=beta, engine=eng)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"M={M} N={N} K={K} eng={eng}: gpu {e0.elapsed_time(e1)*1e3/n:.1f} us/launch, host issue {(t1-t0)*1e6/n:.1f} us, wall {(t2-t0)*1e6/n:.1f} us")
