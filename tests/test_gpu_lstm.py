"""Persistent tcgen05 cluster LSTM kernels (csrc/lstm_cluster.cu: forward and BPTT; csrc/lstm_persistent.cu: the earlier
TF32 forward) against the exact-fp32 stepwise engine and the oracle."""
import pytest
import torch

from deer_b200 import _lib, ops
from helpers import assert_close, cosine
from oracle import deer_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_layer(In, H, seed):
    g = torch.Generator().manual_seed(seed)
    a = (6.0 / (In + 4 * H)) ** 0.5
    b = (6.0 / (H + 4 * H)) ** 0.5
    ws = []
    for _ in range(2):
        ws += [(torch.rand(4 * H, In, generator=g) * 2 - 1) * a, (torch.rand(4 * H, H, generator=g) * 2 - 1) * b,
               (torch.rand(4 * H, generator=g) * 2 - 1) * 0.1, (torch.rand(4 * H, generator=g) * 2 - 1) * 0.1]
    return ws


def run(x, ws, engine, grad):
    ops.set_lstm_engine(engine)
    xs = x.to(DEV).requires_grad_(grad)
    wd = [w.to(DEV).requires_grad_(grad) for w in ws]
    with torch.set_grad_enabled(grad):
        h = ops.bilstm_layer(ops.to_time_major(xs), *wd)
    if grad:
        g = torch.Generator().manual_seed(5)
        pr = torch.randn(h.shape, generator=g).to(DEV)
        (h * pr).sum().backward()
        return h.detach(), [xs.grad] + [w.grad for w in wd]
    return h, None


ENGINE_V1 = 4  # DEER_LSTM_PERSISTENT_V1


@pytest.fixture(autouse=True)
def _restore_lstm_options():
    yield
    _lib.set_option(2, 1)   # DEER_OPT_LSTM_TS
    _lib.set_option(3, 0)   # DEER_OPT_LSTM_TILE
    ops.set_gemm_engine(ops.ENGINE_AUTO)


@pytest.mark.parametrize("B,T,In", [(1, 1, 84), (4, 2, 84), (4, 6, 84), (64, 20, 84), (70, 9, 512), (130, 33, 84)])
@pytest.mark.parametrize("grad", [False, True])
@pytest.mark.parametrize("ts,tile", [(1, 0), (0, 16), (1, 32), (0, 32)])
def test_cluster_matches_stepwise(B, T, In, grad, ts, tile):
    """ts: resident weights in TMEM (1) or shared memory (0); tile: batch columns per cluster (0 = auto)."""
    H = 256
    ws = make_layer(In, H, B + T)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B))
    ops.set_gemm_engine(ops.ENGINE_SIMT)     # exact input projection for both runs: isolates the recurrence
    h_ref, g_ref = run(x, ws, ops.ENGINE_SIMT, grad)
    _lib.set_option(2, ts)
    _lib.set_option(3, tile)
    h_per, g_per = run(x, ws, ops.ENGINE_AUTO, grad)
    assert_close(h_per, h_ref, 1e-3, "h")
    assert float((h_per - h_ref).abs().max()) < 2e-3
    if grad:
        for a, b_ in zip(g_per, g_ref):
            if float(b_.abs().max()) == 0.0:     # T == 1: no recurrent-weight gradient
                assert float(a.abs().max()) == 0.0
            else:
                assert cosine(a, b_) > 0.9999


@pytest.mark.parametrize("B,T,In", [(64, 20, 84), (70, 9, 512)])
def test_persistent_v1_matches_stepwise(B, T, In):
    H = 256
    ws = make_layer(In, H, B + T)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B))
    ops.set_gemm_engine(ops.ENGINE_SIMT)
    h_ref, _ = run(x, ws, ops.ENGINE_SIMT, False)
    h_per, _ = run(x, ws, ENGINE_V1, False)
    assert_close(h_per, h_ref, 1e-3, "h")


def test_cluster_bptt_full_length_vs_oracle():
    """T=300 recurrence, forward (FP16 operands) and BPTT (BF16 operands) against fp64 autograd of the oracle."""
    B, T, In, H = 6, 300, 84, 256
    ws = make_layer(In, H, 99)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(2))
    h, grads = run(x, ws, ops.ENGINE_AUTO, True)
    xd = x.double().requires_grad_(True)
    wd = [w.double().requires_grad_(True) for w in ws]
    f = O.lstm_direction(xd, *wd[:4], False)
    r = O.lstm_direction(xd, *wd[4:], True)
    ref = torch.cat([f, r], dim=-1).permute(1, 0, 2)
    assert_close(h, ref, 1e-3, "h vs oracle")
    pr = torch.randn(h.shape, generator=torch.Generator().manual_seed(5)).double()
    (ref * pr).sum().backward()
    for got, want in zip(grads, [xd.grad] + [w.grad for w in wd]):
        assert cosine(got, want) > 0.999


def test_persistent_full_length_vs_oracle():
    B, T, In, H = 8, 300, 84, 256
    ws = make_layer(In, H, 77)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(1))
    h, _ = run(x, ws, ops.ENGINE_AUTO, False)
    wd = [w.double() for w in ws]
    f = O.lstm_direction(x.double(), *wd[:4], False)
    r = O.lstm_direction(x.double(), *wd[4:], True)
    ref = torch.cat([f, r], dim=-1).permute(1, 0, 2)
    assert_close(h, ref, 1e-3, "h vs oracle")


@pytest.mark.gpu
def test_fused_interlayer_dropout_matches_separate_kernel():
    """nn.LSTM inter-layer dropout fused into the layer-1 operand casts (and into dx) draws the same Philox mask as the
    stand-alone dropout kernel: identical outputs and gradients."""
    import deer_b200
    from deer_b200 import ops
    torch.manual_seed(0)
    enc = deer_b200.EnhancedAudioEncoder({"dropout": 0.3}).cuda().train()
    x = torch.randn(32, 24, 84, device="cuda")
    res = []
    for fused in (True, False):
        ops.set_fuse_lstm_dropout(fused)
        try:
            ops.manual_seed(123)
            ops.begin_step()
            for p in enc.parameters():
                p.grad = None
            xi = x.clone().requires_grad_(True)
            h = enc.lstm_forward(xi)
            (h * torch.linspace(-1, 1, h.numel(), device="cuda").view_as(h)).sum().backward()
            res.append((h.detach().clone(), xi.grad.clone(), enc.lstm.weight_ih_l1.grad.clone(),
                        enc.lstm.weight_hh_l0.grad.clone()))
        finally:
            ops.set_fuse_lstm_dropout(True)
    for a, b in zip(*res):
        assert torch.equal(a, b) or float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
    # and dropout is really applied: eval-mode output differs
    enc.eval()
    with torch.no_grad():
        h_eval = enc.lstm_forward(x)
    assert float((h_eval - res[0][0]).abs().max()) > 1e-3


@pytest.mark.parametrize("B,T,In", [(64, 20, 84), (130, 33, 512), (256, 40, 512)])
def test_fp16_preactivations_match_fp32_preactivations(B, T, In):
    """Input projection written as FP16 by the GEMM epilogue and read as FP16 by the recurrence kernel (default) vs the
    fp32 pre-activation buffer: same h within the FP16-operand noise, same gradients; BPTT writes only the BF16 dpre."""
    H = 256
    ws = make_layer(In, H, B + T + 1)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B + 3))
    res = []
    for on in (True, False):
        ops.set_lstm_pre16(on)
        try:
            res.append(run(x, ws, ops.ENGINE_AUTO, True))
        finally:
            ops.set_lstm_pre16(True)
    (h16, g16), (h32, g32) = res
    assert_close(h16, h32, 3e-4, "h (fp16 pre vs fp32 pre)")
    assert float((h16 - h32).abs().max()) < 2e-3
    for a, b_ in zip(g16, g32):
        assert cosine(a, b_) > 0.99999
    # against the exact engine as well
    ops.set_gemm_engine(ops.ENGINE_SIMT)
    h_ref, g_ref = run(x, ws, ops.ENGINE_SIMT, True)
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    assert_close(h16, h_ref, 1e-3, "h (fp16 pre vs exact fp32)")
    for a, b_ in zip(g16, g_ref):
        assert cosine(a, b_) > 0.9999


@pytest.mark.parametrize("B,T,In", [(70, 9, 512), (130, 33, 84), (1024, 12, 84)])
def test_dual_subtile_forward_matches_monolithic_tile(B, T, In):
    """Inference forward at 32 batch columns per CTA: two interleaved 16-column sub-tiles (DEER_OPT_LSTM_DUAL, default)
    against the monolithic 32-column tile and the exact-fp32 stepwise engine; ragged batches leave a sub-tile partly or
    completely empty."""
    H = 256
    ws = make_layer(In, H, B + T + 7)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B + 1))
    _lib.set_option(3, 32)          # force 32 columns per CTA also for small batches
    try:
        _lib.set_option(15, 0)      # same input path for both: the projection GEMM (the monolithic tile has no other)
        _lib.set_option(9, 1)
        h_dual, _ = run(x, ws, ops.ENGINE_AUTO, False)
        _lib.set_option(9, 0)
        h_mono, _ = run(x, ws, ops.ENGINE_AUTO, False)
        _lib.set_option(15, 1)      # ... and the dual kernel with the input projection inside (In <= 128): FP16-rounding apart
        _lib.set_option(9, 1)
        h_dual_xin, _ = run(x, ws, ops.ENGINE_AUTO, False)
    finally:
        _lib.set_option(15, 1)
        _lib.set_option(9, 1)
        _lib.set_option(3, 0)
    assert torch.equal(h_dual, h_mono) or float((h_dual - h_mono).abs().max()) < 1e-6
    assert_close(h_dual_xin, h_mono, 1e-3, "dual kernel with in-kernel projection")
    if B <= 130:
        ops.set_gemm_engine(ops.ENGINE_SIMT)
        h_ref, _ = run(x, ws, ops.ENGINE_SIMT, False)
        assert_close(h_dual, h_ref, 1e-3, "h vs exact fp32")


@pytest.mark.parametrize("B,T,In", [(4, 6, 84), (70, 9, 512), (130, 33, 84), (256, 40, 512)])
@pytest.mark.parametrize("grad", [False, True])
def test_column_split_forward_matches_eight_warp_kernel(B, T, In, grad):
    """Forward on 16-column tiles with 16 compute warps (two column halves, DEER_OPT_LSTM_COLSPLIT, default) against the
    8-warp kernel: same h, and -- through the unchanged BPTT kernel reading the kept gates / cell states -- the same
    gradients (the kept layouts must be identical)."""
    H = 256
    ws = make_layer(In, H, B + T + 11)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B + 5))
    _lib.set_option(3, 16)
    _lib.set_option(15, 0)       # both runs through the projection GEMM (the in-kernel projection has no column-split form)
    res = []
    try:
        for cs in (1, 0):
            _lib.set_option(10, cs)
            res.append(run(x, ws, ops.ENGINE_AUTO, grad))
    finally:
        _lib.set_option(10, 0)   # the library default
        _lib.set_option(15, 1)
        _lib.set_option(3, 0)
    (h1, g1), (h0, g0) = res
    assert torch.equal(h1, h0) or float((h1 - h0).abs().max()) < 1e-6
    if grad:
        for a, b_ in zip(g1, g0):
            assert_close(a, b_, 1e-5, "gradients")
