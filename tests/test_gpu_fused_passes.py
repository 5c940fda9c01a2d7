"""Elementwise passes folded into their neighbours (round 2): the fused and the separate form must agree -- bit for bit
where both run the same arithmetic on the same Philox stream, to fp32 round-off elsewhere."""
import pytest
import torch

import deer_b200  # noqa: F401
from deer_b200 import ops
from deer_b200.encoders import EnhancedTextEncoder, EnhancedVideoEncoder

from helpers import assert_close, cosine

pytestmark = pytest.mark.gpu


def cu(t):
    return t.detach().clone().float().cuda()


@pytest.mark.parametrize("B,T,binary", [(32, 64, True), (40, 48, False), (3, 10, True)])
def test_text_mask_inside_scorer_and_pooling_matches_materialised_mask(B, T, binary):
    """`token_embeddings * attention_mask` (encoders.py:733-735) applied inside the scorer's operand cast and the pooling
    kernels (scorer_pool(premask=True)) == the materialised masked copy (ops.rowscale) followed by the same node: forward
    and every parameter gradient, for binary masks and for arbitrary float masks (exact semantics, not a 0/1 shortcut);
    (3, 10) is below the split-precision engine's size and takes the materialising fallback inside scorer_pool."""
    g = torch.Generator().manual_seed(B * 100 + T)
    x = torch.randn(B, T, 768, generator=g)
    m = (torch.rand(B, T, generator=g) > 0.3).float() if binary else torch.rand(B, T, generator=g)
    m[:, 0] = 1.0
    if B > 1:
        m[1, 1:] = 0.0              # a row with a single valid token
    pr = torch.randn(B, 768, generator=g)
    enc = EnhancedTextEncoder({"dropout": 0.0}).cuda().train()
    att = enc.token_attention
    res = []
    for premask in (True, False):
        for p in att.parameters():
            p.grad = None
        xc, mc = cu(x), cu(m)
        if premask:
            out, wts = ops.scorer_pool(xc, att[0].weight, att[0].bias, att[2].weight, att[2].bias, mc, premask=True)
        else:
            out, wts = ops.scorer_pool(ops.rowscale(xc, mc), att[0].weight, att[0].bias, att[2].weight, att[2].bias, mc)
        (out * cu(pr)).sum().backward()
        res.append((out.detach(), wts.detach(), [p.grad.clone() for p in att.parameters()]))
    assert_close(res[0][0], res[1][0], 1e-6, "pooled")
    assert_close(res[0][1], res[1][1], 1e-6, "weights")
    for a, b, (n, _) in zip(res[0][2], res[1][2], att.named_parameters()):
        if n == "2.bias":      # the bias in front of a softmax: an exactly-zero derivative, round-off only
            assert float(a.abs().max()) < 1e-4 and float(b.abs().max()) < 1e-4, n
        else:
            assert_close(a, b, 2e-5, n)
    # the encoder takes the fused form by itself and agrees with the fp64 statement of the reference lines
    y = enc(cu(x), cu(m))
    xm = (x * m[..., None]).double()
    w1, b1, w2, b2 = (p.detach().double().cpu() for p in (att[0].weight, att[0].bias, att[2].weight, att[2].bias))
    s = torch.tanh(xm @ w1.T + b1) @ w2.T + b2
    a = torch.softmax(s, dim=1) * m[..., None].double()
    a = a / (a.sum(1, keepdim=True) + 1e-10)
    agg = (xm * a).sum(1)
    assert_close(res[0][0], agg, 1e-4, "pooled vs fp64")
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("B,T,C", [(32, 50, 512), (8, 40, 512)])
def test_conv_dropout_inside_padding_passes_is_bitwise_the_separate_node(B, T, C):
    """nn.Dropout -> nn.Conv1d (encoders.py:453-459) with the dropout folded into the convolution's padding / un-padding
    passes draws the SAME Philox masks as the separate dropout node: outputs identical, gradients identical up to the
    accumulation order of the gradient GEMMs, same zero pattern in dx; (32, 50)
    runs the split-precision forward (A operand written by the fused pass), (8, 40) the TF32 window path."""
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, C, generator=g)
    w = torch.randn(C, C, 3, generator=g) * (3 * C) ** -0.5
    b = torch.randn(C, generator=g) * 0.1
    pr = torch.randn(B, T, C, generator=g)
    res = []
    for fused in (True, False):
        ops.set_conv_dropout_fused(fused)
        try:
            ops.manual_seed(77)
            ops.begin_step()
            xc, wc, bc = (cu(t).requires_grad_(True) for t in (x, w, b))
            pre = ops.dropout(xc * 1.0, 0.25, True)            # an earlier dropout: the offsets must keep their order
            y = ops.conv1d_k3(pre, wc, bc, 0.3, True)
            z = ops.dropout(y, 0.1, True)                       # ... and a later one
            (z * cu(pr)).sum().backward()
            res.append((y.detach(), z.detach(), xc.grad, wc.grad, bc.grad))
        finally:
            ops.set_conv_dropout_fused(True)
    for a, c, n in zip(res[0], res[1], ("y", "z", "dx", "dw", "db")):
        if n in ("y", "z"):
            assert torch.equal(a, c), n
        else:   # the gradient GEMMs accumulate through TMA reduce-adds / split-K in a run-dependent order
            assert_close(a, c, 1e-5, n)
    assert torch.equal(res[0][2] == 0, res[1][2] == 0), "dropout mask of dx"
    zero = float((res[0][2] == 0).float().mean())
    assert 0.35 < zero < 0.6, zero      # 1 - 0.75 * 0.7 = 0.475 of the input gradient is masked


def test_video_encoder_train_mode_dropout_statistics():
    """The video encoder's own use of the fused form: train-mode forward / backward run, are finite, differ between two
    steps (fresh masks) and agree in expectation with eval mode to the dropout noise level."""
    torch.manual_seed(0)
    enc = EnhancedVideoEncoder({"dropout": 0.3, "frame_feature_dim": 256}).cuda().train()
    x = torch.randn(32, 50, 256, device="cuda")
    ops.begin_step()
    y1 = enc(x)
    y1.sum().backward()
    ops.begin_step()
    y2 = enc(x)
    assert torch.isfinite(y1).all() and torch.isfinite(y2).all()
    assert not torch.equal(y1, y2)
    for p in enc.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_lstm_batch_first_input_cast_is_bitwise_the_permuted_copy():
    """First nn.LSTM layer on the batch_first input (encoders.py:82-89,380): one permute + FP16 / BF16 cast pass ==
    the fp32 time-major copy followed by the cast passes (same roundings of the same values): h and all weight gradients
    are identical, in train mode (BF16 copy used by dW_ih) and in eval mode."""
    from deer_b200.encoders import EnhancedAudioEncoder
    torch.manual_seed(3)
    enc = EnhancedAudioEncoder({"dropout": 0.0}).cuda().train()
    x = torch.randn(6, 37, 84, device="cuda")
    pr = torch.randn(37, 6, 512, device="cuda")
    res = []
    for fused in (True, False):
        ops.set_lstm_batch_major_input(fused)
        try:
            for p in enc.parameters():
                p.grad = None
            h = enc.lstm_forward(x)
            (h * pr).sum().backward()
            with torch.no_grad():
                he = enc.lstm_forward(x)
            res.append((h.detach(), he, [p.grad.clone() for p in enc.lstm.parameters()]))
        finally:
            ops.set_lstm_batch_major_input(True)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for a, b, (n, _) in zip(res[0][2], res[1][2], enc.lstm.named_parameters()):
        assert torch.equal(a, b), n


def test_lstm_interlayer_dropout_mask_in_gemm_epilogue_matches_philox_pass():
    """nn.LSTM inter-layer dropout (encoders.py:82-89) in backward: the keep bits written by the forward's dropout + cast
    pass, applied by the epilogue of layer 1's input-gradient GEMM (deer_gemm_h16_dropmask) == the Philox pass over dx:
    same forward, the same gradients (identical operands, same kernel: bitwise)."""
    from deer_b200.encoders import EnhancedAudioEncoder
    torch.manual_seed(5)
    enc = EnhancedAudioEncoder({"dropout": 0.3}).cuda().train()
    x = torch.randn(8, 40, 84, device="cuda")           # M = 320 rows: the CTA-pair kernel
    pr = torch.randn(40, 8, 512, device="cuda")
    res = []
    for masked in (True, False):
        ops.set_lstm_dropout_mask(masked)
        try:
            ops.manual_seed(11)
            ops.begin_step()
            for p in enc.parameters():
                p.grad = None
            h = enc.lstm_forward(x)
            (h * pr).sum().backward()
            res.append((h.detach(), {n: p.grad.clone() for n, p in enc.lstm.named_parameters()}))
        finally:
            ops.set_lstm_dropout_mask(True)
    assert torch.equal(res[0][0], res[1][0])
    for n in res[0][1]:
        a, b = res[0][1][n], res[1][1][n]
        if "_l1" in n:        # layer 1's own gradients do not pass through the masked dx
            assert torch.equal(a, b), n
        else:                 # layer 0 sees dx: x * scale vs x * (1/(1-p)) may differ in the last bit
            assert_close(a, b, 1e-6, n)


def test_dropout_keep_mask_bits_match_the_dropped_tensor():
    from deer_b200._lib import call
    n = 128 * 50
    x = torch.randn(n, device="cuda").abs() + 0.1
    y16 = torch.empty(n, device="cuda", dtype=torch.float16)
    mask = torch.zeros(n // 32, device="cuda", dtype=torch.int32)
    call("deer_dropout_cast16", x.data_ptr(), y16.data_ptr(), None, n, 0.4, 99, 7, None, mask.data_ptr())
    bits = ((mask.view(-1, 1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1).reshape(-1).bool()
    assert torch.equal(bits, y16 != 0)
    assert 0.5 < float(bits.float().mean()) < 0.7


@pytest.mark.parametrize("B,T,train", [(6, 37, True), (32, 20, True), (256, 12, True), (300, 9, False), (320, 7, False),
                                       (40, 1, True)])
def test_lstm_input_projection_inside_the_recurrence_matches_the_gemm_path(B, T, train):
    """First nn.LSTM layer (encoders.py:82-89,380; In = 84): W_ih x_t issued inside the forward recurrence kernel
    (deer_lstm_cluster_fwd_xin: 16-column tiles for one-wave batches, dual sub-tiles with one MMA warp each for the no-keep
    forward of larger ones; whole-tile and ragged batches) == the projection GEMM + FP16 pre-activations path up to that path's own FP16
    rounding of the pre-activations, and both agree with torch's fp64 nn.LSTM."""
    from deer_b200 import _lib
    from deer_b200.encoders import EnhancedAudioEncoder
    torch.manual_seed(B + T)
    enc = EnhancedAudioEncoder({"dropout": 0.0}).cuda()
    enc.train(train)
    x = torch.randn(B, T, 84, device="cuda")
    pr = torch.randn(T, B, 512, device="cuda")
    assert _lib.load().deer_lstm_cluster_xin_mode(B, int(train), 88) == (1 if B <= 256 else 2)
    res = []
    for fused in (True, False):
        ops.set_lstm_input_projection_fused(fused)
        try:
            for p in enc.parameters():
                p.grad = None
            if train:
                h = enc.lstm_forward(x)
                (h * pr).sum().backward()
                res.append((h.detach(), [p.grad.clone() for p in enc.lstm.parameters()]))
            else:
                with torch.no_grad():
                    res.append((enc.lstm_forward(x), []))
        finally:
            ops.set_lstm_input_projection_fused(True)
    assert_close(res[0][0], res[1][0], 1e-3, "h")
    for a, b, (n, _) in zip(res[0][1], res[1][1], enc.lstm.named_parameters()):
        # (two BF16-operand backward passes over pre-activations that differ by one FP16 rounding: the north-star gate)
        if float(b.abs().max()) == 0.0:      # T = 1: no recurrent step, dW_hh is exactly zero
            assert float(a.abs().max()) == 0.0, n
            continue
        assert_close(a, b, 1e-2, n)
        assert cosine(a, b) >= 0.9999, n
    ref = torch.nn.LSTM(84, 256, 2, batch_first=True, bidirectional=True).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in enc.lstm.state_dict().items()})
    hr, _ = ref(x.double().cpu())
    e_fused = float((res[0][0].double().cpu().permute(1, 0, 2) - hr).norm() / hr.norm())
    e_gemm = float((res[1][0].double().cpu().permute(1, 0, 2) - hr).norm() / hr.norm())
    assert e_fused < 1e-3 and e_fused < 1.5 * e_gemm + 1e-5, (e_fused, e_gemm)


@pytest.mark.parametrize("B,T,time_major", [(64, 150, True), (200, 50, False), (33, 300, True)])
def test_pooling_input_gradient_inside_the_scorer_gemm_epilogue(B, T, time_major):
    """Backward of scorer + attention pooling (encoders.py:93-98,383-384): dx = w[b,t] dout[b,:] + dh W1 written ONCE by the
    scorer's input-gradient GEMM (deer_gemm_rowterm) == the pooling kernel writing its part and the GEMM accumulating onto
    it: same dx (fp32 addition order apart) and identical parameter gradients; time-major (audio) and batch-major (video)."""
    D = 512
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn((T, B, D) if time_major else (B, T, D), generator=g)
    pr = torch.randn(B, D, generator=g)
    enc = EnhancedVideoEncoder({"dropout": 0.0}).cuda().train()
    att = enc.temporal_attention
    res = []
    for on in (True, False):
        ops.set_pool_rowterm(on)
        try:
            for p in att.parameters():
                p.grad = None
            xc = cu(x).requires_grad_(True)
            out, _ = ops.scorer_pool(xc, att[0].weight, att[0].bias, att[2].weight, att[2].bias, None, time_major,
                                     precise=False)
            (out * cu(pr)).sum().backward()
            res.append((out.detach(), xc.grad.clone(), [p.grad.clone() for p in att.parameters()]))
        finally:
            ops.set_pool_rowterm(True)
    assert torch.equal(res[0][0], res[1][0])
    assert_close(res[0][1], res[1][1], 1e-6, "dx")
    for a, b, (n, _) in zip(res[0][2], res[1][2], att.named_parameters()):
        if n == "2.bias":
            assert float(a.abs().max()) < 1e-4 and float(b.abs().max()) < 1e-4, n
        else:
            assert_close(a, b, 1e-5, n)
    # fp64 statement of the same lines
    xd = x.double().requires_grad_(True)
    w1, b1, w2, b2 = (p.detach().double().cpu() for p in (att[0].weight, att[0].bias, att[2].weight, att[2].bias))
    xb = xd.permute(1, 0, 2) if time_major else xd
    s = torch.tanh(xb @ w1.T + b1) @ w2.T + b2
    o = (torch.softmax(s, dim=1) * xb).sum(1)
    (o * pr.double()).sum().backward()
    assert_close(res[0][1], xd.grad, 2e-3, "dx vs fp64")
