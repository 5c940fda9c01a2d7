"""The post-pooling chain of the sequence model -- HierarchicalMultimodalFusion.forward (fusion.py:119-171) followed by
MultiDimensionalDEER's evidence network (deer.py:233-266) -- as ONE persistent-kernel launch per direction
(csrc/chain.cu, deer_chain_run) instead of ~50 / ~70 dependent launches.

`fusion_head_chain(fusion, deer, a, v, t)` returns the same tensors the module-by-module path produces
(fused / audiovisual / trimodal features, head-averaged 2-token attention weights, raw evidence [B,D,4]); the module
path (fusion.py / deer.py on the fused 3xTF32 nodes of ops.py) remains the reference it is tested against and the
fallback for shapes the chain kernel does not cover.

Program layout (B = batch rows, E1 = intermediate width, E = fusion width, H = head width):
  forward  18 levels / 26 ops   L1 {vp, ap, tp}  L2 u = [vp;ap] Wv  L3 z = u Wo  L4 y1 = relu([a_att|v_att] Wf)  L5 LN
           L6 avp  L7 qkv = [avp;tp] Win  L8 2-token attention  L9 out_proj  L10 y2  L11 LN  L12 y3  L13 LN
           L14-15 feature_processor  L16-18 the D heads
  backward the mirror image, every layer's dW (+ bias-gradient column sums) in the level of its dx.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib, ops
from ._lib import GemmX3Args, I, P

MAX_OPS, MAX_LEVELS = 72, 48
K_GEMM, K_LN_FWD, K_LN_BWD, K_MHA2_FWD, K_MHA2_BWD, K_AXPY = range(6)


class ChainOp(ctypes.Structure):
    _fields_ = [("kind", I), ("tiles", I), ("tiles_n", I), ("pad_", I), ("g", GemmX3Args)]


class ChainProgram(ctypes.Structure):
    _fields_ = [("nops", I), ("nlevels", I), ("barrier", P), ("error", P), ("level_begin", I * (MAX_LEVELS + 1)),
                ("ops", ChainOp * MAX_OPS)]


def _p(t):
    """Device address of a tensor / (tensor, element offset) pair / int / None."""
    if t is None:
        return None
    if isinstance(t, tuple):
        return t[0].data_ptr() + 4 * int(t[1])
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    return int(t)


class Program:
    """Builder of a deer_chain_program: ops appended to the current level; `level()` opens the next one."""

    def __init__(self, device):
        self.c = ChainProgram()
        self.n = 0
        self.levels = [0]
        self.device = device

    def level(self):
        if self.n > self.levels[-1]:
            self.levels.append(self.n)

    def _op(self, kind, tiles) -> GemmX3Args:
        if self.n >= MAX_OPS:
            raise _lib.DeerError("deer_b200.chain: program too long")
        op = self.c.ops[self.n]
        op.kind, op.tiles = kind, int(tiles)
        self.n += 1
        return op.g

    def gemm(self, A, lda, transA, B, ldb, transB, C, ldc, M, N, K, *, bias=None, act=0, beta=0.0, gate=None, ldgate=0,
             gate_mode=0, gate_scale=1.0, colsum=None, drop=None, drop_ld=0, drop_col0=0):
        g = self._op(K_GEMM, ((M + 31) // 32) * ((N + 31) // 32))
        g.A, g.B, g.C, g.bias, g.gate, g.colsum = _p(A), _p(B), _p(C), _p(bias), _p(gate), _p(colsum)
        g.lda, g.ldb, g.ldc, g.ldgate = int(lda), int(ldb), int(ldc), int(ldgate)
        g.M, g.N, g.K, g.batch = int(M), int(N), int(K), 1
        g.transA, g.transB, g.act, g.beta = int(transA), int(transB), int(act), float(beta)
        g.gate_mode, g.gate_scale = int(gate_mode), float(gate_scale)
        if drop is not None and drop[0] > 0.0:
            g.drop_p, g.drop_seed, g.drop_offset, g.drop_step = float(drop[0]), int(drop[1]), int(drop[2]), _p(drop[3])
            g.drop_ld, g.drop_col0 = int(drop_ld), int(drop_col0)

    def ln_fwd(self, x, ldx, gamma, beta, y, ldy, mean, rstd, M, N, eps):
        g = self._op(K_LN_FWD, (M + 7) // 8)
        g.A, g.B, g.bias, g.C, g.gate, g.colsum = _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd)
        g.lda, g.ldc, g.M, g.N, g.beta = int(ldx), int(ldy), int(M), int(N), float(eps)

    def ln_bwd(self, dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, N, accumulate=False):
        g = self._op(K_LN_BWD, (M + 31) // 32)
        g.A, g.B, g.bias, g.gate, g.colsum, g.C = _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx)
        g.lda, g.ldb, g.ldc, g.M, g.N = int(lddy), int(ldx), int(lddx), int(M), int(N)
        g.beta = 1.0 if accumulate else 0.0
        g.drop_seed, g.drop_offset = _p(dgamma) or 0, _p(dbeta) or 0

    def mha2_fwd(self, qkv, ctx_mean, ldc, attw, probs, B, E, heads):
        g = self._op(K_MHA2_FWD, B)
        g.A, g.C, g.colsum, g.bias = _p(qkv), _p(ctx_mean), _p(attw), _p(probs)
        g.ldc, g.M, g.N, g.K = int(ldc), int(B), int(E), int(heads)

    def mha2_bwd(self, qkv, dctx_mean, ld, dattw, probs, dqkv, B, E, heads):
        g = self._op(K_MHA2_BWD, B)
        g.A, g.B, g.gate, g.bias, g.C = _p(qkv), _p(dctx_mean), _p(dattw), _p(probs), _p(dqkv)
        g.ldb, g.M, g.N, g.K = int(ld), int(B), int(E), int(heads)

    def axpy(self, x, ldx, y, ldy, M, N, accumulate=False):
        g = self._op(K_AXPY, (M + 31) // 32)
        g.A, g.C, g.lda, g.ldc, g.M, g.N = _p(x), _p(y), int(ldx), int(ldy), int(M), int(N)
        g.beta = 1.0 if accumulate else 0.0

    def run(self):
        self.level()
        nl = len(self.levels) - 1
        if self.levels[-1] != self.n:
            nl += 1
            self.levels.append(self.n)
        if nl > MAX_LEVELS:
            raise _lib.DeerError("deer_b200.chain: too many levels")
        c = self.c
        c.nops, c.nlevels = self.n, nl
        for i, b in enumerate(self.levels):
            c.level_begin[i] = b
        scratch = ops.zeros_scratch(16, self.device)      # [0]: grid-barrier counter, [8]: watchdog flag
        c.barrier = scratch.data_ptr()
        c.error = scratch.data_ptr() + 32
        _lib.check(_lib.load().deer_chain_run(ctypes.byref(c), _lib.stream()), "deer_chain_run")


# OFF by default: measured on one box, back to back (tools/gpu_r2_ab.sh, B = 256 training step, graph replay): 4.171 ms
# with the chain kernel vs 4.112 ms module by module (the fused 3xTF32 nodes, weight gradients on their own stream).  The
# kernel replaces ~66 launches per step, but inside a CUDA graph a launch costs ~2 us while every level of the chain pays
# a grid-wide barrier plus a cold operand pipeline (~10 us; profiles/r2_chain_rowpartition_experiment.txt).
_state = {"enabled": False, "max_batch": 512}


def set_max_batch(n: int):
    """Largest batch that takes the chain kernel (default 512).  Measured (tools/chain_time.py, graph replay): the kernel
    replaces ~120 launches of 5-8 us but keeps one grid barrier + one cold operand pipeline per level, so it only pays
    while the layers are launch/latency-bound; from B = 1024 on the module-by-module path is faster (373 vs 510 us
    forward at B = 1024, 651 vs 957 us at B = 2048)."""
    _state["max_batch"] = int(n)


def set_enabled(on: bool):
    """Fusion + NIG head as one persistent-kernel launch per direction, or module by module (default)."""
    _state["enabled"] = bool(on)


def enabled() -> bool:
    return _state["enabled"] and ops._fused_chain()


def _params(fusion, deer):
    av, tri = fusion.audio_visual_fusion, fusion.trimodal_fusion
    ca, ma = av.cross_attention, tri.modality_attention
    fl, ff, op, fp = av.fusion_layers, tri.final_fusion, fusion.output_projection, deer.feature_processor
    ps = [av.audio_projection.weight, av.audio_projection.bias, av.video_projection.weight, av.video_projection.bias,
          ca.in_proj_weight, ca.in_proj_bias, ca.out_proj.weight, ca.out_proj.bias, fl[0].weight, fl[0].bias,
          fl[3].weight, fl[3].bias,
          tri.audiovisual_projection.weight, tri.audiovisual_projection.bias, tri.text_projection.weight,
          tri.text_projection.bias, ma.in_proj_weight, ma.in_proj_bias, ma.out_proj.weight, ma.out_proj.bias,
          ff[0].weight, ff[0].bias, ff[3].weight, ff[3].bias,
          op[0].weight, op[0].bias, op[3].weight, op[3].bias,
          fp[0].weight, fp[0].bias, fp[3].weight, fp[3].bias]
    for h in deer.deer_heads:
        n = h.evidence_net
        ps += [n[0].weight, n[0].bias, n[3].weight, n[3].bias, n[6].weight, n[6].bias]
    return ps


def supported(fusion, deer, a, v, t) -> bool:
    try:
        ps = _params(fusion, deer)
    except AttributeError:
        return False
    E1, E = ps[0].shape[0], ps[12].shape[0]
    dims = [a.shape[1], v.shape[1], t.shape[1], E1, E, ps[28].shape[0], ps[32].shape[0], ps[34].shape[0]]
    drops = {float(fusion.dropout), float(fusion.audio_visual_fusion.dropout), float(fusion.trimodal_fusion.dropout),
             float(deer.dropout)} | {float(h.dropout) for h in deer.deer_heads}
    return (a.dim() == 2 and a.shape[0] <= _state["max_batch"] and len(drops) == 1 and
            all(d % 4 == 0 for d in dims) and E <= 512 and E1 <= 512 and
            fusion.training == deer.training and
            tri_heads(fusion) <= 8 and E % tri_heads(fusion) == 0 and ps[36].shape[0] == 4 and
            all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps) and
            len(deer.deer_heads) * (6) + 32 == len(ps))


def tri_heads(fusion) -> int:
    return int(fusion.trimodal_fusion.num_heads)


class _FusionHeadChain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, heads, p_drop, training, eps, a, v, t, *ps):
        a, v, t = (ops._req(x, "features").contiguous() for x in (a, v, t))
        dev = a.device
        B = a.shape[0]
        Da, Dv, Dt = a.shape[1], v.shape[1], t.shape[1]
        E1, E, H = ps[0].shape[0], ps[12].shape[0], ps[28].shape[0]
        D = (len(ps) - 32) // 6
        H1, H2 = ps[32].shape[0], ps[34].shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        new = lambda *s: torch.empty(s, **f32)   # noqa: E731
        pv, u, z = new(B, 2, E1), new(B, 2, E1), new(B, 2, E1)
        y1, av, at = new(B, E1), new(B, E1), new(B, 2, E)
        qkv, cm, probs, attw = new(B, 2, 3 * E), new(B, E), new(B, heads, 4), new(B, 2, 2)
        pooled, y2, tri, y3, fused = new(B, E), new(B, E), new(B, E), new(B, E), new(B, E)
        f1, f2, h1, h2, ev = new(B, H), new(B, H), new(B, D, H1), new(B, D, H2), new(B, D, 4)
        stats = new(6, B)    # mean / rstd of the three LayerNorms
        # dropout reservations in the order of the module path (identical masks): y1, y2, y3, f1, f2, h1, h2
        dr = [ops._take_dropout(n, p_drop, training) for n in (B * E1, B * E, B * E, B * H, B * H, B * D * H1, B * D * H2)]
        P_ = Program(dev)
        g = P_.gemm
        g(v, Dv, 0, ps[2], Dv, 1, pv, 2 * E1, B, E1, Dv, bias=ps[3])                      # vp -> pv[:,0]
        g(a, Da, 0, ps[0], Da, 1, (pv, E1), 2 * E1, B, E1, Da, bias=ps[1])                # ap -> pv[:,1]
        g(t, Dt, 0, ps[14], Dt, 1, (at, E), 2 * E, B, E, Dt, bias=ps[15])                 # tp -> at[:,1]
        P_.level()
        g(pv, E1, 0, (ps[4], 2 * E1 * E1), E1, 1, u, E1, 2 * B, E1, E1, bias=(ps[5], 2 * E1))   # V rows of in_proj
        P_.level()
        g(u, E1, 0, ps[6], E1, 1, z, E1, 2 * B, E1, E1, bias=ps[7])                       # z[b] = [a_att | v_att]
        P_.level()
        g(z, 2 * E1, 0, ps[8], 2 * E1, 1, y1, E1, B, E1, 2 * E1, bias=ps[9], act=1, drop=dr[0])
        P_.level()
        P_.ln_fwd(y1, E1, ps[10], ps[11], av, E1, stats[0], stats[1], B, E1, eps[0])
        P_.level()
        g(av, E1, 0, ps[12], E1, 1, at, 2 * E, B, E, E1, bias=ps[13])                     # avp -> at[:,0]
        P_.level()
        g(at, E, 0, ps[16], E, 1, qkv, 3 * E, 2 * B, 3 * E, E, bias=ps[17])
        P_.level()
        P_.mha2_fwd(qkv, cm, E, attw, probs, B, E, heads)
        P_.level()
        g(cm, E, 0, ps[18], E, 1, pooled, E, B, E, E, bias=ps[19])
        P_.level()
        g(pooled, E, 0, ps[20], E, 1, y2, E, B, E, E, bias=ps[21], act=1, drop=dr[1])
        P_.level()
        P_.ln_fwd(y2, E, ps[22], ps[23], tri, E, stats[2], stats[3], B, E, eps[1])
        P_.level()
        g(tri, E, 0, ps[24], E, 1, y3, E, B, E, E, bias=ps[25], act=1, drop=dr[2])
        P_.level()
        P_.ln_fwd(y3, E, ps[26], ps[27], fused, E, stats[4], stats[5], B, E, eps[2])
        P_.level()
        g(fused, E, 0, ps[28], E, 1, f1, H, B, H, E, bias=ps[29], act=1, drop=dr[3])
        P_.level()
        g(f1, H, 0, ps[30], H, 1, f2, H, B, H, H, bias=ps[31], act=1, drop=dr[4])
        P_.level()
        for d in range(D):
            g(f2, H, 0, ps[32 + 6 * d], H, 1, (h1, d * H1), D * H1, B, H1, H, bias=ps[33 + 6 * d], act=1, drop=dr[5],
              drop_ld=D * H1, drop_col0=d * H1)
        P_.level()
        for d in range(D):
            g((h1, d * H1), D * H1, 0, ps[34 + 6 * d], H1, 1, (h2, d * H2), D * H2, B, H2, H1, bias=ps[35 + 6 * d], act=1,
              drop=dr[6], drop_ld=D * H2, drop_col0=d * H2)
        P_.level()
        for d in range(D):
            g((h2, d * H2), D * H2, 0, ps[36 + 6 * d], H2, 1, (ev, 4 * d), 4 * D, B, 4, H2, bias=ps[37 + 6 * d])
        P_.run()
        ctx.save_for_backward(a, v, t, pv, u, z, y1, av, at, qkv, cm, probs, pooled, y2, tri, y3, fused, f1, f2, h1, h2,
                              stats, *ps)
        ctx.set_materialize_grads(False)     # unused outputs arrive as None in backward, not as zero tensors
        ctx.cfg = (heads, float(p_drop) if (training and p_drop > 0.0) else 0.0)
        ctx.params = ps
        ctx.mark_non_differentiable(attw)
        return fused, av, tri, attw, ev

    @staticmethod
    def backward(ctx, dfused_up, dav_up, dtri_up, _dattw, dev):
        sv = ctx.saved_tensors
        (a, v, t, pv, u, z, y1, av, at, qkv, cm, probs, pooled, y2, tri, y3, fused, f1, f2, h1, h2, stats) = sv[:22]
        ps = sv[22:]
        pp = ctx.params
        heads, p_drop = ctx.cfg
        dev_ = a.device
        B = a.shape[0]
        Da, Dv, Dt = a.shape[1], v.shape[1], t.shape[1]
        E1, E, H = ps[0].shape[0], ps[12].shape[0], ps[28].shape[0]
        D = (len(ps) - 32) // 6
        H1, H2 = ps[32].shape[0], ps[34].shape[0]
        gs = 1.0 / (1.0 - p_drop) if p_drop > 0.0 else 1.0
        f32 = dict(device=dev_, dtype=torch.float32)
        new = lambda *s: torch.empty(s, **f32)   # noqa: E731
        if dev is None:
            dev = torch.zeros((B, D, 4), **f32)
        dev = dev.contiguous()
        # gradient accumulators of the parameters (param.grad itself in trainer mode)
        acc = [ops._acc(p) if ctx.needs_input_grad[7 + i] else (None, False) for i, p in enumerate(pp)]
        G = [x[0] for x in acc]
        need = [ctx.needs_input_grad[4 + i] for i in range(3)]
        dh2, dh1, df2p, df2, df1 = new(B, D, H2), new(B, D, H1), new(D, B, H), new(B, H), new(B, H)
        dfu, dy3, dtri, dy2, dpooled, dcm = new(B, E), new(B, E), new(B, E), new(B, E), new(B, E), new(B, E)
        dqkv, dat, dav, dy1 = new(B, 2, 3 * E), new(B, 2, E), new(B, E1), new(B, E1)
        dz, du, dpv = new(B, 2, E1), new(B, 2, E1), new(B, 2, E1)
        da = new(B, Da) if need[0] else None
        dv = new(B, Dv) if need[1] else None
        dt = new(B, Dt) if need[2] else None
        P_ = Program(dev_)
        g = P_.gemm

        def wgrad(dy, lddy, x, ldx, wi, N, K, M, gate=None, ldgate=0, off_w=0, off_b=0):
            """dW[wi] += (dy (.) gate)^T x  (+ bias gradient by column sums into G[wi+1])."""
            if G[wi] is None:
                return
            g(dy, lddy, 1, x, ldx, 0, (G[wi], off_w), K, N, K, M, beta=1.0, gate=gate, ldgate=ldgate,
              gate_mode=1 if gate is not None else 0, gate_scale=gs if gate is not None else 1.0,
              colsum=None if G[wi + 1] is None else (G[wi + 1], off_b))

        # ---- heads
        if dfused_up is not None:
            P_.axpy(dfused_up.contiguous(), E, dfu, E, B, E)
        if dtri_up is not None:
            P_.axpy(dtri_up.contiguous(), E, dtri, E, B, E)
        if dav_up is not None:
            P_.axpy(dav_up.contiguous(), E1, dav, E1, B, E1)
        for d in range(D):
            g((dev, 4 * d), 4 * D, 0, ps[36 + 6 * d], H2, 0, (dh2, d * H2), D * H2, B, H2, 4)
            wgrad((dev, 4 * d), 4 * D, (h2, d * H2), D * H2, 36 + 6 * d, 4, H2, B)
        P_.level()
        for d in range(D):
            g((dh2, d * H2), D * H2, 0, ps[34 + 6 * d], H1, 0, (dh1, d * H1), D * H1, B, H1, H2, gate=(h2, d * H2),
              ldgate=D * H2, gate_mode=1, gate_scale=gs)
            wgrad((dh2, d * H2), D * H2, (h1, d * H1), D * H1, 34 + 6 * d, H2, H1, B, gate=(h2, d * H2), ldgate=D * H2)
        P_.level()
        for d in range(D):
            g((dh1, d * H1), D * H1, 0, ps[32 + 6 * d], H, 0, (df2p, d * B * H), H, B, H, H1, gate=(h1, d * H1),
              ldgate=D * H1, gate_mode=1, gate_scale=gs)
            wgrad((dh1, d * H1), D * H1, f2, H, 32 + 6 * d, H1, H, B, gate=(h1, d * H1), ldgate=D * H1)
        P_.level()
        for d in range(D):      # df2 = sum over the heads (they share the input f2)
            P_.axpy((df2p, d * B * H), H, df2, H, B, H, accumulate=d > 0)
            P_.level()
        g(df2, H, 0, ps[30], H, 0, df1, H, B, H, H, gate=f2, ldgate=H, gate_mode=1, gate_scale=gs)
        wgrad(df2, H, f1, H, 30, H, H, B, gate=f2, ldgate=H)
        P_.level()
        g(df1, H, 0, ps[28], E, 0, dfu, E, B, E, H, gate=f1, ldgate=H, gate_mode=1, gate_scale=gs,
          beta=1.0 if dfused_up is not None else 0.0)
        wgrad(df1, H, fused, E, 28, H, E, B, gate=f1, ldgate=H)
        P_.level()
        # ---- output_projection
        P_.ln_bwd(dfu, E, y3, E, ps[26], stats[4], stats[5], dy3, E, G[26], G[27], B, E)
        P_.level()
        g(dy3, E, 0, ps[24], E, 0, dtri, E, B, E, E, gate=y3, ldgate=E, gate_mode=1, gate_scale=gs,
          beta=1.0 if dtri_up is not None else 0.0)
        wgrad(dy3, E, tri, E, 24, E, E, B, gate=y3, ldgate=E)
        P_.level()
        # ---- trimodal fusion
        P_.ln_bwd(dtri, E, y2, E, ps[22], stats[2], stats[3], dy2, E, G[22], G[23], B, E)
        P_.level()
        g(dy2, E, 0, ps[20], E, 0, dpooled, E, B, E, E, gate=y2, ldgate=E, gate_mode=1, gate_scale=gs)
        wgrad(dy2, E, pooled, E, 20, E, E, B, gate=y2, ldgate=E)
        P_.level()
        g(dpooled, E, 0, ps[18], E, 0, dcm, E, B, E, E)
        wgrad(dpooled, E, cm, E, 18, E, E, B)
        P_.level()
        P_.mha2_bwd(qkv, dcm, E, None, probs, dqkv, B, E, heads)
        P_.level()
        g(dqkv, 3 * E, 0, ps[16], E, 0, dat, E, 2 * B, E, 3 * E)
        wgrad(dqkv, 3 * E, at, E, 16, 3 * E, E, 2 * B)
        P_.level()
        g(dat, 2 * E, 0, ps[12], E1, 0, dav, E1, B, E1, E, beta=1.0 if dav_up is not None else 0.0)
        wgrad(dat, 2 * E, av, E1, 12, E, E1, B)
        if dt is not None:
            g((dat, E), 2 * E, 0, ps[14], Dt, 0, dt, Dt, B, Dt, E)
        wgrad((dat, E), 2 * E, t, Dt, 14, E, Dt, B)
        P_.level()
        # ---- audio-visual fusion
        P_.ln_bwd(dav, E1, y1, E1, ps[10], stats[0], stats[1], dy1, E1, G[10], G[11], B, E1)
        P_.level()
        g(dy1, E1, 0, ps[8], 2 * E1, 0, dz, 2 * E1, B, 2 * E1, E1, gate=y1, ldgate=E1, gate_mode=1, gate_scale=gs)
        wgrad(dy1, E1, z, 2 * E1, 8, E1, 2 * E1, B, gate=y1, ldgate=E1)
        P_.level()
        g(dz, E1, 0, ps[6], E1, 0, du, E1, 2 * B, E1, E1)
        wgrad(dz, E1, u, E1, 6, E1, E1, 2 * B)
        P_.level()
        g(du, E1, 0, (ps[4], 2 * E1 * E1), E1, 0, dpv, E1, 2 * B, E1, E1)
        wgrad(du, E1, pv, E1, 4, E1, E1, 2 * B, off_w=2 * E1 * E1, off_b=2 * E1)     # V rows; Q / K rows stay exactly zero
        P_.level()
        if dv is not None:
            g(dpv, 2 * E1, 0, ps[2], Dv, 0, dv, Dv, B, Dv, E1)
        if da is not None:
            g((dpv, E1), 2 * E1, 0, ps[0], Da, 0, da, Da, B, Da, E1)
        wgrad(dpv, 2 * E1, v, Dv, 2, E1, Dv, B)
        wgrad((dpv, E1), 2 * E1, a, Da, 0, E1, Da, B)
        P_.run()
        grads = [None if (buf is None or direct) else buf for buf, direct in acc]
        return (None, None, None, None, da, dv, dt, *grads)


def fusion_head_chain(fusion, deer, a, v, t):
    """-> (fused_features, audiovisual_features, trimodal_features, trimodal attention weights [B,2,2], evidence
    [B,D,4]) of fusion(a, v, t) followed by deer.evidence(fused), one kernel launch (and one in backward)."""
    ps = _params(fusion, deer)
    eps = (fusion.audio_visual_fusion.fusion_layers[3].eps, fusion.trimodal_fusion.final_fusion[3].eps,
           fusion.output_projection[3].eps)
    return _FusionHeadChain.apply(tri_heads(fusion), float(fusion.dropout), bool(fusion.training), eps, a, v, t, *ps)
