"""GPU probe for the persistent cluster LSTM kernels: correctness against the exact-fp32 stepwise engine and timing.

    python tools/lstm_probe.py --ts 1 --tile 16 --B 256 --T 300 [--time]

Each configuration should run in its own process (a device trap poisons the CUDA context)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200  # noqa: E402
from deer_b200 import _lib  # noqa: E402
from deer_b200._lib import call, ptr  # noqa: E402

SIMT, AUTO, V1 = 1, 0, 4


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ts", type=int, default=1)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--T", type=int, default=300)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--engine", type=int, default=AUTO)
    ap.add_argument("--skip-bwd", action="store_true")
    ap.add_argument("--prof", action="store_true")
    ap.add_argument("--dual", type=int, default=1)
    ap.add_argument("--stasync", type=int, default=0)
    ap.add_argument("--halfsplit", type=int, default=0)
    a = ap.parse_args()
    _lib.set_option(9, a.dual)
    _lib.set_option(12, a.stasync)
    _lib.set_option(13, a.halfsplit)
    _lib.set_option(2, a.ts)
    _lib.set_option(3, a.tile)
    dev = "cuda"
    T, B, H = a.T, a.B, 256
    g = torch.Generator().manual_seed(3)
    s = (6.0 / (5 * H)) ** 0.5
    w = [((torch.rand(4 * H, H, generator=g) * 2 - 1) * s).to(dev) for _ in range(2)]
    pre = (torch.randn(T, B, 2, 4 * H, generator=g) * 1.5).to(dev)
    dh_out = (torch.randn(T, B, 2 * H, generator=g) * 1e-3).to(dev)
    tag = f"ts={a.ts} tile={a.tile} B={B} T={T} engine={a.engine}"

    Bp = (B + 31) // 32 * 32
    G = 4 * H

    def il(t):      # [...,4H] natural -> interleaved
        return t.view(*t.shape[:-1], 4, H).transpose(-1, -2).contiguous().view(t.shape)

    def nat(t):     # interleaved -> natural
        return t.view(*t.shape[:-1], H, 4).transpose(-1, -2).contiguous().view(t.shape)

    def fwd(engine):
        if engine == AUTO:
            pre_il = il(pre)
            h = torch.empty(T, B, 2 * H, device=dev)
            gact = torch.empty(T * 2 * Bp * G, device=dev)
            c = torch.empty(T * 2 * Bp * H, device=dev)
            call("deer_lstm_cluster_fwd", ptr(pre_il), ptr(w[0]), ptr(w[1]), ptr(h), ptr(gact), ptr(c), None, None, T, B, H)
            torch.cuda.synchronize()
            return gact, h, c
        gates = pre.clone()
        h = torch.empty(T, B, 2 * H, device=dev)
        c = torch.empty(T, B, 2, H, device=dev)
        call("deer_lstm_fwd", ptr(gates), ptr(w[0]), ptr(w[1]), ptr(h), ptr(c), None, T, B, H, engine)
        torch.cuda.synchronize()
        return gates, h, c

    def bwd(engine, gates, c):
        if engine == AUTO:
            dpre = torch.empty(T, B, 2, G, device=dev)
            db = torch.zeros(2, G, device=dev)
            call("deer_lstm_cluster_bwd", ptr(gates), ptr(c), ptr(dh_out), ptr(w[0]), ptr(w[1]), ptr(dpre), ptr(db), None, T, B, H)
            torch.cuda.synchronize()
            return nat(dpre), nat(db)
        gt = gates.clone()
        dhw = torch.empty(B, 2, H, device=dev)
        dcw = torch.empty(B, 2, H, device=dev)
        call("deer_lstm_bwd", ptr(gt), ptr(w[0]), ptr(w[1]), ptr(c), ptr(dh_out), ptr(dhw), ptr(dcw), T, B, H, engine)
        torch.cuda.synchronize()
        return gt, gt.sum(dim=(0, 1))

    g_ref, h_ref, c_ref = fwd(SIMT)
    g_new, h_new, c_new = fwd(a.engine)
    print(f"[{tag}] fwd: h rel={rel(h_new, h_ref):.3e} max={float((h_new - h_ref).abs().max()):.3e}", flush=True)
    # per-time-step error growth (first bad step localises protocol bugs)
    per_t = ((h_new - h_ref).abs().amax(dim=(1, 2))).cpu()
    bad = (per_t > 5e-3).nonzero().flatten()
    print(f"[{tag}] fwd: first bad t (fwd dir view) = {bad[:4].tolist() if len(bad) else None}; "
          f"err t0={float(per_t[0]):.2e} t1={float(per_t[min(1, T - 1)]):.2e} tmid={float(per_t[T // 2]):.2e}", flush=True)
    if not a.skip_bwd:
        d_ref, db_ref = bwd(SIMT, g_ref, c_ref)
        d_new, db_new = bwd(a.engine, g_new, c_new)
        print(f"[{tag}] bwd: dgates rel={rel(d_new, d_ref):.3e} cos={cos(d_new, d_ref):.7f} db rel={rel(db_new, db_ref):.3e}",
              flush=True)
        pt = ((d_new - d_ref).flatten(1).norm(dim=1) / d_ref.flatten(1).norm(dim=1).clamp_min(1e-30)).cpu()
        print(f"[{tag}] bwd: per-t rel err t=T-1 {float(pt[-1]):.2e} T-2 {float(pt[-2]):.2e} mid {float(pt[T // 2]):.2e} "
              f"t=0 {float(pt[0]):.2e}", flush=True)
    if a.prof:
        lib = _lib.load()
        for what in ("fwd", "bwd"):
            buf = torch.zeros(32, dtype=torch.int64, device=dev)
            lib.deer_lstm_set_profile_buffer(buf.data_ptr())
            if what == "fwd":
                fwd(a.engine)
            else:
                bwd(a.engine, g_new, c_new)
            lib.deer_lstm_set_profile_buffer(None)
            v = buf.cpu().view(4, 8)
            for i in range(4):
                base = int(v[i, 0])
                nxt = int(v[i + 1, 0]) - base if i + 1 < 4 else -1
                print(f"[{tag}] prof {what} step {64 + i}: " + " ".join(f"s{k}={int(v[i, k]) - base:+d}" for k in range(1, 6))
                      + f" | next step s0 at +{nxt}", flush=True)
    if a.time:
        def timeit(fn, n=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        pre_il = il(pre)
        h = torch.empty(T, B, 2 * H, device=dev)
        tf = timeit(lambda: call("deer_lstm_cluster_fwd", ptr(pre_il), ptr(w[0]), ptr(w[1]), ptr(h), ptr(g_new), ptr(c_new), None, None, T, B, H))
        ti = timeit(lambda: call("deer_lstm_cluster_fwd", ptr(pre_il), ptr(w[0]), ptr(w[1]), ptr(h), None, None, None, None, T, B, H))
        msg = f"[{tag}] time: fwd(keep) {tf:.3f} ms = {tf * 1e3 / T:.2f} us/step; fwd(infer) {ti:.3f} ms"
        if not a.skip_bwd:
            dpre = torch.empty(T, B, 2, G, device=dev)
            db = torch.zeros(2, G, device=dev)
            tb = timeit(lambda: call("deer_lstm_cluster_bwd", ptr(g_new), ptr(c_new), ptr(dh_out), ptr(w[0]), ptr(w[1]),
                                     ptr(dpre), ptr(db), None, T, B, H))
            msg += f"; bwd {tb:.3f} ms = {tb * 1e3 / T:.2f} us/step"
        print(msg, flush=True)


if __name__ == "__main__":
    main()
