"""Time the exact-fp32 small-problem GEMM engines on the post-pooling shapes of a B=256 training step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deer_b200
from deer_b200 import ops, _lib

SHAPES = [  # M, N, K, ta, tb   (forward y = x W^T; dX = dY W; dW = dY^T X)
    (256, 512, 512, 0, 1), (256, 512, 512, 0, 0), (512, 512, 256, 1, 0),
    (256, 256, 256, 0, 1), (256, 256, 512, 0, 1), (256, 1536, 512, 0, 1), (256, 128, 256, 0, 1),
    (256, 512, 640, 0, 1), (640, 512, 256, 1, 0), (256, 64, 128, 0, 1), (256, 4, 64, 0, 1),
]


def main():
    dev = torch.device("cuda")
    for eng in (1, 2):
        _lib.set_option(4, eng)
        tot = 0.0
        for (M, N, K, ta, tb) in SHAPES:
            A = torch.randn((K, M) if ta else (M, K), device=dev)
            B = torch.randn((N, K) if tb else (K, N), device=dev)
            C = torch.empty(M, N, device=dev)
            bias = torch.randn(N, device=dev)
            f = lambda: ops.gemm(A, A.shape[1], ta, B, B.shape[1], tb, C, N, M, N, K, bias=bias, act=1,
                                 engine=ops.ENGINE_SIMT)
            for _ in range(3):
                f()
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                f()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(20):
                        f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 200
            tot += us
            print(f"engine {eng}  M={M:4d} N={N:4d} K={K:4d} ta={ta} tb={tb}: {us:7.2f} us  "
                  f"{2.0*M*N*K/us/1e6:6.2f} TFLOP/s", flush=True)
        print(f"engine {eng} total {tot:.1f} us")


if __name__ == "__main__":
    main()
