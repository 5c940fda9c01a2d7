"""Time the two DEER-loss passes (deer_nig_loss_stats / deer_nig_loss_finish) alone, per size, with operands rotating
through enough buffers that every iteration starts with its inputs in HBM.  python tools/nig_probe.py [log2 sizes...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deer_b200
from deer_b200 import ops
from deer_b200._lib import call, ptr

PEAK = 6547.2


def events(fn, n, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    dev = torch.device("cuda")
    sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 21, 22]
    edges = ops.ece_edges(dev)
    for lg in sizes:
        n = 1 << lg
        D = 3
        per_buf = n * D * 64
        nbuf = max(2, min(8, int(300e6 // per_buf) + 1))
        ev = [torch.randn(n, D, 4, device=dev) for _ in range(nbuf)]
        tg = [torch.tanh(torch.randn(n, D, device=dev)) for _ in range(nbuf)]
        nig = [torch.empty(7, n, D, device=dev) for _ in range(nbuf)]
        grad = [torch.empty(n, D, 4, device=dev) for _ in range(nbuf)]
        stats = torch.zeros(D, 40, device=dev)
        losses = torch.empty(5 * D + 2, device=dev)

        def p1(i):
            j = i % nbuf
            call("deer_nig_loss_stats", ptr(ev[j]), None, None, None, None, ptr(tg[j]), ptr(edges), ptr(stats),
                 ptr(nig[j]), n, D, 1, 1e-8)

        def p2(i):
            j = i % nbuf
            call("deer_nig_loss_finish", ptr(ev[j]), None, None, None, None, ptr(tg[j]), ptr(edges), ptr(stats), None,
                 0.1, 0.01, 0.05, 0.05, 1e-8, n, n, D, 1, 1.0, ptr(losses), ptr(grad[j]))

        def both(i):
            stats.zero_()
            p1(i)
            p2(i)

        def api(i):
            j = i % nbuf
            ops.nig_loss_raw(ev[j], None, tg[j], want_nig=True, want_grad=True)

        for pipe in (1, 0):   # DEER_OPT_NIG_PIPELINE
            deer_b200._lib.set_option(8, pipe)
            t1, t2, tb, ta = events(p1, 20), events(p2, 20), events(both, 20), events(api, 20)
            alg = n * D * 64.0
            print(f"B=2^{lg} ({nbuf} bufs, pipe={pipe}): stats {t1:8.1f} us ({n*D*48/t1/1e3:7.0f} GB/s)  finish {t2:8.1f} us "
                  f"({n*D*36/t2/1e3:7.0f} GB/s)  both {tb:8.1f} us  api {ta:8.1f} us -> algorithmic "
                  f"{alg/tb/1e3:7.0f} GB/s = {alg/tb/1e3/PEAK*100:5.1f}% (api {alg/ta/1e3/PEAK*100:5.1f}%)", flush=True)
        deer_b200._lib.set_option(8, 1)
        del ev, tg, nig, grad
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
