"""GPU probe: time of the fusion + NIG-head chain (forward, forward + backward) with the persistent grid-barrier chain
kernel and the module-by-module path, at the benchmark batch sizes (CUDA events, L2 flushed between
replays by a 256 MB fill so the weights come from HBM as they do inside a step; the step is captured into a CUDA graph and
timed by deer_timestamp marks, so no host time is included)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import deer_b200  # noqa: E402
from deer_b200 import chain, ops  # noqa: E402
from deer_b200.deer import nig_dict  # noqa: E402

DEV = "cuda"


def run(B, mode, train, iters=10):
    chain.set_enabled(mode != "modules")
    chain.set_max_batch(1 << 30)
    torch.manual_seed(0)
    fus = deer_b200.HierarchicalMultimodalFusion(512, 512, 512, fusion_dim=512, intermediate_dim=256,
                                                 dropout=0.3 if train else 0.0).to(DEV)
    head = deer_b200.MultiDimensionalDEER(512, 3, 256, 0.3 if train else 0.0).to(DEV)
    fus.train(train), head.train(train)
    g = torch.Generator().manual_seed(B)
    a, v, t = (torch.randn(B, 512, generator=g).to(DEV).requires_grad_(train) for _ in range(3))
    flush = torch.empty(64 << 20, device=DEV)

    def step():
        flush.fill_(1.0)
        ops.begin_step()
        with torch.set_grad_enabled(train):
            ops.mark("t0")
            if mode == "modules":
                f = fus(a, v, t)
                ev = head.evidence(f["fused_features"])
            else:
                _, _, _, _, ev = chain.fusion_head_chain(fus, head, a, v, t)
            ops.mark("t1")
            if train:
                ev.backward(torch.ones_like(ev))
                ops.mark("t2")            # dx chain done on the main stream (the encoders' backward could start here)
                ops.join_wgrad_stream()
                ops.mark("t3")            # weight gradients done

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.synchronize()
    ops.timeline_begin(DEV)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            step()
    tf = tb = tw = 0.0
    for it in range(iters):
        graph.replay()
        torch.cuda.synchronize()
        tl = dict(ops.timeline_read())
        tf += tl["t1"] - tl["t0"]
        if train:
            tb += tl["t2"] - tl["t1"]
            tw += tl["t3"] - tl["t1"]
    ops.timeline_end()
    return tf / iters / 1e3, tb / iters / 1e3, tw / iters / 1e3


for B, train in ((256, True), (512, True), (1024, False), (64, True), (2048, True)):
    for mode in ("grid", "modules"):
        f, b, w = run(B, mode, train)
        print(f"B={B:5d} {'train' if train else 'infer'} {mode:8s}: forward {f:8.1f} us   backward (dx chain) {b:8.1f} us"
              f"   backward incl. weight gradients {w:8.1f} us", flush=True)

