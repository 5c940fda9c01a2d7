"""GPU parity: the CUDA path (through the C ABI) against (1) golden vectors produced by the unmodified reference and
(2) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): NIG parameters and loss within 1e-3 relative, gradient cosine >= 0.999.
The fp32 engines are held to a tighter TOL_FP32; the tcgen05/TF32 engines to the north-star 1e-3."""
import numpy as np
import pytest
import torch

import deer_b200
from deer_b200 import ops
from gen_common import det_normal, nig_inputs, probe, seq_inputs
from helpers import assert_close, check_param_grads, cosine, load_fixture_weights, rel_l2
from oracle import deer_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
TOL_FP32 = 2e-4
DEV = "cuda"


def cu(t):
    return t.to(torch.float32).to(DEV)


@pytest.fixture(autouse=True)
def _engine():
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    yield


# ----------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("M,N,K,ta,tb", [(5, 7, 3, 0, 1), (130, 70, 84, 0, 1), (64, 64, 64, 1, 0), (33, 129, 65, 0, 0),
                                         (200, 4, 64, 0, 1), (96, 100, 4100, 1, 0), (17, 31, 29, 1, 1),
                                         # >= 148 64x64 tiles: the large-tile kernel (the rest use the 32x32 split-K one)
                                         (1100, 900, 70, 0, 1), (900, 1100, 33, 1, 0), (256, 512, 512, 0, 1),
                                         (512, 256, 256, 1, 0), (256, 768, 640, 0, 0),
                                         # cp.async-pipelined small kernel: every operand orientation with M/N/K tails
                                         (36, 52, 100, 1, 1), (100, 36, 68, 0, 0), (44, 28, 200, 1, 0),
                                         (130, 72, 132, 0, 1), (256, 384, 256, 0, 1), (256, 256, 1536, 0, 0)])
def test_gemm_simt(M, N, K, ta, tb):
    g = torch.Generator().manual_seed(M * 131 + N)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
    for act, beta in ((0, 0.0), (1, 0.0), (2, 1.0), (0, 1.0)):
        C = cu(C0).clone()
        ops.gemm(cu(A), A.shape[1], ta, cu(B), B.shape[1], tb, C, N, M, N, K, bias=cu(bias), act=act, beta=beta,
                 engine=ops.ENGINE_SIMT)
        r = ref + bias.double() + beta * C0.double()
        r = {0: r, 1: torch.relu(r), 2: torch.tanh(r)}[act]
        assert_close(C, r, 1e-5, f"gemm act={act} beta={beta}")


def test_layernorm_and_linear_autograd():
    ops.set_gemm_engine(ops.ENGINE_SIMT)   # op wiring test: exact engine, tight tolerance
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 96, generator=g)
    w = torch.randn(128, 96, generator=g) * 0.1
    b = torch.randn(128, generator=g) * 0.1
    lg, lb = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    pr = torch.randn(37, 128, generator=g)

    def run(fl, fn, args):
        args = [a.clone().requires_grad_(True) for a in args]
        y = fn(*args)
        (y * pr.to(y)).sum().backward()
        return y, [a.grad for a in args]
    yr, gr = run(None, lambda x, w, b, lg, lb: O.layer_norm(torch.relu(O.linear(x, w, b)), lg, lb),
                 [t.double() for t in (x, w, b, lg, lb)])
    yc, gc = run(None, lambda x, w, b, lg, lb: ops.layer_norm(ops.linear(x, w, b, "relu"), lg, lb),
                 [cu(t) for t in (x, w, b, lg, lb)])
    assert_close(yc, yr, TOL_FP32, "y")
    for a, r, n in zip(gc, gr, "x w b lg lb".split()):
        assert_close(a, r, TOL_FP32, "grad " + n)


def test_linear_concat_inputs_matches_cat():
    ops.set_gemm_engine(ops.ENGINE_SIMT)
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(9, 40, generator=g), torch.randn(9, 24, generator=g)
    w, bias = torch.randn(32, 64, generator=g) * 0.2, torch.randn(32, generator=g)
    ad, bd, wd = a.double().requires_grad_(True), b.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = torch.tanh(O.linear(torch.cat([ad, bd], 1), wd, bias.double()))
    yr.sum().backward()
    ac, bc, wc = cu(a).requires_grad_(True), cu(b).requires_grad_(True), cu(w).requires_grad_(True)
    yc = ops.linear([ac, bc], wc, cu(bias), "tanh")
    yc.sum().backward()
    assert_close(yc, yr, TOL_FP32)
    assert_close(ac.grad, ad.grad, TOL_FP32)
    assert_close(bc.grad, bd.grad, TOL_FP32)
    assert_close(wc.grad, wd.grad, TOL_FP32)


def test_dropout_statistics_and_backward_mask():
    x = torch.ones(1 << 20, device=DEV, requires_grad=True)
    ops.manual_seed(123)
    y = ops.dropout(x, 0.3, True)
    keep = float((y != 0).float().mean())
    assert abs(keep - 0.7) < 5e-3
    assert abs(float(y.mean()) - 1.0) < 1e-2
    y.sum().backward()
    assert torch.equal((x.grad != 0), (y != 0))
    ops.manual_seed(123)
    y2 = ops.dropout(torch.ones(1 << 20, device=DEV), 0.3, True)
    assert torch.equal(y2, y.detach())
    assert ops.dropout(x, 0.3, False) is x


# ----------------------------------------------------------------------------- modules vs golden fixtures
@pytest.mark.parametrize("name", ["audio_small", "audio_full_t12"])
def test_audio_encoder_golden(golden, name):
    fx = golden(name)
    m = fx.meta
    enc = load_fixture_weights(deer_b200.EnhancedAudioEncoder({"hidden_dim": m["hidden"], "dropout": 0.0}), fx).to(DEV)
    x = cu(seq_inputs(m["B"], m["T"], 2, 2, seed=m["seed"])[0]).requires_grad_(True)
    # H=256 runs the persistent cluster kernels and the 16-bit tcgen05 GEMM engine (FP16 operands forward, BF16 in
    # BPTT and the weight gradients, fp32 accumulation): held to the north-star tolerance, not the fp32-engine one
    fast = m["hidden"] == 512
    otol = TOL if fast else TOL_FP32
    h_tm = enc.lstm_forward(x)
    assert_close(h_tm.permute(1, 0, 2), fx.t("lstm_out"), otol, "lstm_out")
    y = enc(x)
    assert_close(y, fx.t("out"), otol, "out")
    (y * cu(probe("audio_out", y.shape, m["seed"]))).sum().backward()
    gtol = 5 * TOL if fast else TOL_FP32     # BF16 operands: ~2^-9 per element; the gate on gradients is cosine >= 0.999
    assert_close(x.grad, fx.t("dx"), gtol, "dx")
    assert cosine(x.grad, fx.t("dx")) > 0.99999
    check_param_grads(enc, fx, m["seed"], gtol)


@pytest.mark.parametrize("name", ["video_small_train", "video_small_eval", "video_small_f1"])
def test_video_encoder_golden(golden, name):
    fx = golden(name)
    m = fx.meta
    enc = deer_b200.EnhancedVideoEncoder({"hidden_dim": m["hidden"], "dropout": 0.0, "frame_feature_dim": m["din"]})
    enc = load_fixture_weights(enc, fx).to(DEV)
    enc.train(m["training"])
    x = cu(seq_inputs(m["B"], 2, m["F"], 2, Dv=m["din"], seed=m["seed"])[1]).requires_grad_(True)
    y = enc(x)
    assert_close(y, fx.t("out"), TOL_FP32, "out")
    (y * cu(probe("video_out", y.shape, m["seed"]))).sum().backward()
    assert_close(x.grad, fx.t("dx"), TOL_FP32, "dx")
    check_param_grads(enc, fx, m["seed"], TOL_FP32)
    if m["training"] and m["F"] > 1:
        sd = enc.state_dict()
        for k in fx.keys("post:"):
            assert_close(sd[k[5:]], fx.t(k), 1e-5, k)
        assert int(sd["temporal_cnn.1.num_batches_tracked"]) == 1


def test_text_encoder_golden(golden):
    fx = golden("text_small")
    m = fx.meta
    enc = load_fixture_weights(deer_b200.EnhancedTextEncoder({"hidden_dim": m["hidden"], "dropout": 0.0}), fx).to(DEV)
    _, _, tok, mask, ling, _ = seq_inputs(m["B"], 2, 2, m["T"], seed=m["seed"])
    tok, ling = cu(tok).requires_grad_(True), cu(ling).requires_grad_(True)
    y = enc(tok, cu(mask), ling)
    assert_close(y, fx.t("out"), TOL_FP32, "out")
    (y * cu(probe("text_out", y.shape, m["seed"]))).sum().backward()
    assert_close(tok.grad, fx.t("dtok"), TOL_FP32, "dtok")
    assert_close(ling.grad, fx.t("dling"), TOL_FP32, "dling")
    check_param_grads(enc, fx, m["seed"], TOL_FP32)


def test_fusion_golden(golden):
    fx = golden("fusion_small")
    da, dv, dt, fd, idim, heads = fx.meta["dims"]
    B, seed = fx.meta["B"], fx.meta["seed"]
    fus = deer_b200.HierarchicalMultimodalFusion(da, dv, dt, fusion_dim=fd, intermediate_dim=idim,
                                                 num_attention_heads=heads, dropout=0.0)
    fus = load_fixture_weights(fus, fx).to(DEV)
    a = cu(det_normal("in:fa", (B, da), seed)).requires_grad_(True)
    v = cu(det_normal("in:fv", (B, dv), seed)).requires_grad_(True)
    t = cu(det_normal("in:ft", (B, dt), seed)).requires_grad_(True)
    out = fus(a, v, t)
    keys = ("fused_features", "audiovisual_features", "trimodal_features", "trimodal_attention_weights")
    for k in keys:
        assert_close(out[k], fx.t(k), TOL_FP32, k)
    assert_close(out["av_attention_weights"]["audio_to_video"], fx.t("a2v"), 1e-6)
    assert_close(out["av_attention_weights"]["video_to_audio"], fx.t("v2a"), 1e-6)
    assert out["uncertainty_weights"] is None
    sum((out[k] * cu(probe(k, out[k].shape, seed))).sum() for k in keys).backward()
    assert_close(a.grad, fx.t("da"), TOL_FP32, "da")
    assert_close(v.grad, fx.t("dv"), TOL_FP32, "dv")
    assert_close(t.grad, fx.t("dt"), TOL_FP32, "dt")
    check_param_grads(fus, fx, seed, TOL_FP32)
    # Q/K rows of the seq-1 cross attention get exactly zero gradient; uncertainty_gate.* is never touched
    E = idim
    gw = fus.audio_visual_fusion.cross_attention.in_proj_weight.grad
    assert float(gw[:2 * E].abs().max()) == 0.0
    for n, p in fus.named_parameters():
        if n.startswith("uncertainty_gate"):
            assert p.grad is None


def test_head_golden(golden):
    fx = golden("head_small")
    m = fx.meta
    head = load_fixture_weights(deer_b200.MultiDimensionalDEER(m["din"], 3, m["hidden"], 0.0), fx).to(DEV)
    x = cu(det_normal("in:hx", (m["B"], m["din"]), m["seed"])).requires_grad_(True)
    out = head(x)
    ref_keys = [k for k in fx.arrays if not k.startswith(("grad:", "gsum:", "ghead:")) and k != "dx"]
    assert len(ref_keys) == 23
    for k in ref_keys:
        assert_close(out[k], fx.t(k), TOL_FP32, k)
    sum((out[k] * cu(probe(k, out[k].shape, m["seed"]))).sum() for k in ref_keys).backward()
    assert_close(x.grad, fx.t("dx"), TOL_FP32, "dx")
    check_param_grads(head, fx, m["seed"], TOL_FP32)


@pytest.mark.parametrize("name", ["loss_b64", "loss_b1000", "loss_b3"])
@pytest.mark.parametrize("fused", [True, False])
def test_losses_golden(golden, name, fused):
    fx = golden(name)
    e, y = nig_inputs(fx.meta["B"], fx.meta["seed"])
    e, y = cu(e).requires_grad_(True), cu(y)
    nig = ops.nig_head(e)
    from deer_b200.deer import nig_dict
    pred = nig_dict(e, nig, ["valence", "arousal", "dominance"])
    if not fused:
        pred.pop("_deer_evidence")
    out = deer_b200.MultiTaskDEERLoss()(pred, y)
    for k in fx.keys("mt:"):
        if k == "mt:devidence":
            continue
        ref = float(fx.arrays[k])
        got = float(out[k[3:]])
        assert abs(got - ref) <= TOL_FP32 * max(abs(ref), 1e-3) * 5, (k, got, ref)
    out["total_loss"].backward()
    assert_close(e.grad, fx.t("mt:devidence"), TOL_FP32 * 5, "devidence")
    assert cosine(e.grad, fx.t("mt:devidence")) > 0.99999
    # single-dimension DEERLoss (L1) and the Amini-style deer.DEERLoss (L3)
    single = deer_b200.DEERLoss()({"mu": nig[0][:, 0:1], "nu": nig[1][:, 0:1], "alpha": nig[2][:, 0:1],
                                   "beta": nig[3][:, 0:1]}, y[:, 0:1])
    for k in fx.keys("l1:"):
        ref, got = float(fx.arrays[k]), float(single[k[3:]])
        assert abs(got - ref) <= 1e-3 * max(abs(ref), 1e-3), (k, got, ref)
    e.grad = None
    nig2 = ops.nig_head(e)
    am = deer_b200.AminiDEERLoss()({"mu": nig2[0], "nu": nig2[1], "alpha": nig2[2], "beta": nig2[3]}, y)
    for k in fx.keys("l3:"):
        if k != "l3:devidence":
            ref, got = float(fx.arrays[k]), float(am[k[3:]])
            assert abs(got - ref) <= 1e-3 * max(abs(ref), 1e-3), (k, got, ref)
    am["total_loss"].backward()
    assert_close(e.grad, fx.t("l3:devidence"), 1e-3, "l3 devidence")


# ----------------------------------------------------------------------------- the full sequence composite
def _run_composite(model, batch, engine):
    ops.set_gemm_engine(engine)
    audio, video, text, mask, ling, y = batch
    model.zero_grad(set_to_none=True)
    out = model(cu(audio), cu(video), cu(text), cu(mask), cu(ling))
    loss = model.compute_loss(out, cu(y))
    loss["total_loss"].backward()
    return out, loss


def test_sequence_composite_golden_and_oracle(golden):
    fx = golden("seq_full_b4")
    m = fx.meta
    model = load_fixture_weights(deer_b200.SequenceDEERModel(dropout=0.0), fx).to(DEV)
    model.train()
    batch = seq_inputs(m["B"], m["Ta"], m["Tv"], m["Tt"], seed=m["seed"])
    out, loss = _run_composite(model, batch, ops.ENGINE_AUTO)
    for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features"):
        assert_close(out[k], fx.t(k), TOL, k)
    for k in fx.keys("out:"):
        assert_close(out[k[4:]], fx.t(k), TOL, k)
    for k in fx.keys("loss:"):
        if k.endswith("batch_size"):
            continue
        ref, got = float(fx.arrays[k]), float(loss[k[5:]])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-2), (k, got, ref)
    check_param_grads(model, fx, m["seed"], 5 * TOL)
    # full gradient cosine against the oracle (float64 CPU) on the same inputs
    sd = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float64 else v) for k, v in fx.state_dict().items()}
    _, oloss = O.sequence_model_loss(*batch[:5], batch[5], sd, training=True)
    oloss["total_loss"].backward()
    flat_c, flat_o = [], []
    for n, p in model.named_parameters():
        og = sd[n].grad
        if og is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        if float(og.abs().max()) < 1e-12:   # exactly-zero derivative up to round-off (bias before a softmax, ...)
            assert float(p.grad.abs().max()) < 1e-5, n
            continue
        c = cosine(p.grad, og)
        assert c >= 0.999, (n, c)
        flat_c.append(p.grad.flatten().double().cpu())
        flat_o.append(og.flatten())
    assert cosine(torch.cat(flat_c), torch.cat(flat_o)) >= 0.999


def test_composite_ragged_and_eval_vs_oracle(golden):
    """eval() mode (BatchNorm running statistics), ragged text masks, batch that is not a tile multiple."""
    fx = golden("seq_full_b4")
    model = load_fixture_weights(deer_b200.SequenceDEERModel(dropout=0.0), fx).to(DEV).eval()
    batch = seq_inputs(7, 33, 9, 17, seed=99)
    with torch.no_grad():
        out = model(*[cu(t) for t in batch[:5]])
    sd = fx.state_dict()
    ref = O.sequence_model(*batch[:5], sd, training=False)
    for k in ("mu_all", "uncertainty_all", "valence_nu", "arousal_alpha", "dominance_beta", "fused_features"):
        assert_close(out[k], ref[k], TOL, k)


def test_trainer_direct_grad_accumulation_matches_autograd():
    """The trainer lets the backward kernels accumulate into the flat gradient buffer (ops.set_direct_grad_accumulation);
    the result must equal the autograd-accumulated gradients."""
    import deer_b200
    from deer_b200.trainer import DEERDataParallelTrainer
    from gen_common import seq_inputs

    grads = []
    for direct in (False, True):
        torch.manual_seed(3)
        model = deer_b200.SequenceDEERModel(dropout=0.0).to(DEV).train()
        tr = DEERDataParallelTrainer(model)
        tr.direct_grad = direct
        b = [t.float().to(DEV) for t in seq_inputs(6, 9, 5, 7, seed=11)]
        batch = {"audio_features": b[0], "video_features": b[1], "text_features": b[2], "attention_mask": b[3],
                 "linguistic_features": b[4], "targets": b[5]}
        tr.forward_backward(batch)
        grads.append(tr.flat.grads.clone())
        assert not ops._state["direct_grad"]
    assert float(grads[0].norm()) > 0
    assert_close(grads[1], grads[0], 1e-5, "flat gradient buffer")


def test_pooled_model_golden(golden):
    """CompleteDEERModel (complete_project.py:462-588) + MultiTaskDEERLoss against the reference's own outputs, loss
    components and parameter gradients (tests/golden/pooled_b16.npz, generated from the unmodified reference)."""
    from gen_common import pooled_inputs
    fx = golden("pooled_b16")
    m = fx.meta
    cfg = deer_b200.ModelConfig(dropout=0.0)
    model = load_fixture_weights(deer_b200.CompleteDEERModel(cfg), fx).to(DEV).train()
    for mod in model.modules():      # UncertaintyEstimator has a hard-coded Dropout(0.2) (complete_project.py:193);
        if isinstance(mod, torch.nn.Dropout):   # the golden generator zeroes every dropout the same way
            mod.p = 0.0
    a, v, t, y = pooled_inputs(m["B"], m["seed"])
    out = model(cu(a), cu(v), cu(t))
    for k in fx.keys("out:"):
        assert_close(out[k[4:]], fx.t(k), TOL, k)
    assert out["valence_mu"].shape == (m["B"],) and out["mu_all"].shape == (m["B"], 3)
    loss = model.compute_loss(out, cu(y))
    for k in fx.keys("loss:"):
        if k.endswith("batch_size"):
            continue
        ref, got = float(fx.arrays[k]), float(loss[k[5:]])
        assert abs(got - ref) <= TOL * max(abs(ref), 1e-2), (k, got, ref)
    loss["total_loss"].backward()
    check_param_grads(model, fx, m["seed"], 5 * TOL)
    for k, has in m["has_grad"].items():
        p = dict(model.named_parameters())[k]
        if not has:   # query/key projections (single-key attention), calibration layer: exactly no gradient
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    # dict-argument call used by the reference driver (run_multimodal_deer.py:719) and the uncertainty accessor
    with torch.no_grad():
        out2 = model.eval()({"audio": cu(a), "video": cu(v), "text": cu(t)})
    pred, unc = model.get_predictions_and_uncertainties(out2)
    assert "gamma" in out2 and pred.shape == (m["B"], 3) and unc.shape == (m["B"], 3)
    assert_close(out2["mu_all"], fx.t("out:mu_all"), TOL, "eval mu_all")


def test_pooled_model_train_dropout_runs():
    """Training-mode dropout path (incl. the per-head attention-weight dropout, complete_project.py:172)."""
    torch.manual_seed(0)
    model = deer_b200.CompleteDEERModel(deer_b200.ModelConfig()).to(DEV).train()
    B = 64
    out = model(torch.randn(B, 84, device=DEV), torch.randn(B, 256, device=DEV), torch.randn(B, 768, device=DEV))
    loss = model.compute_loss(out, torch.tanh(torch.randn(B, 3, device=DEV)))
    loss["total_loss"].backward()
    assert torch.isfinite(loss["total_loss"])
    g = model.audio_encoder.input_projection[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0


@pytest.mark.parametrize("B,scale", [(257, 1.5), (4099, 6.0), (65, 12.0)])
def test_fused_loss_wide_range_vs_oracle(B, scale):
    """The fused head+loss kernels use MUFU-based softplus / lgamma / digamma: check them far outside the golden
    fixtures' evidence range (softplus linear branch at x>20, alpha from 1+1e-5 to ~70, tiny nu and beta) against the
    fp32-precision oracle evaluated in float64 (losses.py:72-348)."""
    g = torch.Generator().manual_seed(B)
    e = torch.randn(B, 3, 4, generator=g, dtype=torch.float64) * scale
    # keep alpha - 1 = softplus(e2) >= ~7e-3: (alpha - 1) is formed in fp32 (ulp(1) = 1.2e-7), so below that the
    # reference's own fp32 result is ill-conditioned at the 1e-3 level
    e[:, :, 2] = e[:, :, 2].clamp(min=-5.0)
    y = torch.tanh(torch.randn(B, 3, generator=g, dtype=torch.float64))
    ed = e.clone().requires_grad_(True)
    pred = {}
    for i, d in enumerate(("valence", "arousal", "dominance")):
        for k, v in O.nig_from_evidence(ed[:, i:i + 1, :]).items():
            pred[f"{d}_{k}"] = v
    ref = O.multitask_deer_loss(pred, y)
    ref["total_loss"].backward()
    ec = cu(e).requires_grad_(True)
    nig_out, losses = ops.fused_head_loss(ec, cu(y))
    losses[-1].backward()
    assert abs(float(losses[-1]) - float(ref["total_loss"])) <= TOL * abs(float(ref["total_loss"]))
    for i, d in enumerate(("valence", "arousal", "dominance")):
        for j, k in enumerate(("mu", "nu", "alpha", "beta")):
            assert_close(nig_out[j][:, i], pred[f"{d}_{k}"].detach()[:, 0], 1e-5, f"{d}_{k}")
    assert cosine(ec.grad, ed.grad) > 0.99999
    assert_close(ec.grad, ed.grad, TOL, "d loss / d evidence")


@pytest.mark.parametrize("M,N,K,ta,tb", [(300, 520, 200, 0, 1), (128, 256, 64, 0, 1), (1000, 84, 4100, 1, 0),
                                         (260, 300, 130, 0, 0), (130, 512, 96, 1, 1), (2050, 1024, 84, 0, 1),
                                         # CTA-pair (cta_group::2) kernel over several persistent rounds / split-K
                                         (40000, 520, 512, 0, 1), (1024, 512, 9000, 1, 0), (700, 1024, 1024, 0, 0)])
@pytest.mark.parametrize("a_bf,b_bf", [(0, 0), (1, 1)])
def test_gemm_h16(M, N, K, ta, tb, a_bf, b_bf):
    """16-bit-operand persistent tcgen05 engine: all operand majors, FP16/BF16 mixes, K/M/N tails, bias + activation +
    beta epilogue, split-K accumulation and the 16-bit shadow output, against fp64 on the SAME rounded operands."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    Kp, Mp, Np = (K + 7) // 8 * 8, (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.randn(M, K, generator=g) * 0.5
    B = torch.randn(N, K, generator=g) * 0.5
    da, db_ = (torch.bfloat16 if a_bf else torch.float16), (torch.bfloat16 if b_bf else torch.float16)
    # stored operands (zero-padded to a 16-byte pitch): [M,Kp] or [K,Mp] for A, [N,Kp] or [K,Np] for B
    if ta:
        As = torch.zeros(K, Mp, dtype=da); As[:, :M] = A.t().to(da); lda = Mp
        Ar = As[:, :M].double().t()
    else:
        As = torch.zeros(M, Kp, dtype=da); As[:, :K] = A.to(da); lda = Kp
        Ar = As[:, :K].double()
    if tb:
        Bs = torch.zeros(N, Kp, dtype=db_); Bs[:, :K] = B.to(db_); ldb = Kp
        Br = Bs[:, :K].double()
    else:
        Bs = torch.zeros(K, Np, dtype=db_); Bs[:, :N] = B.t().to(db_); ldb = Np
        Br = Bs[:, :N].double().t()
    ref = Ar @ Br.t()
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    Ad, Bd = As.to(DEV), Bs.to(DEV)
    ldc = (N + 3) // 4 * 4
    for act, beta in ((0, 0.0), (2, 0.0), (0, 1.0), (1, 0.0)):
        C = torch.zeros(M, ldc, device=DEV)
        C[:, :N] = cu(C0)
        use_bias = bias if N % 4 == 0 else None     # the TMA epilogue loads the bias as 16-byte vectors
        ops.gemm_h16(Ad, lda, ta, Bd, ldb, tb, C, ldc, M, N, K, a_bf16=bool(a_bf), b_bf16=bool(b_bf),
                     bias=None if use_bias is None else cu(use_bias), act=act, beta=beta)
        r = ref + (0 if use_bias is None else use_bias.double()) + beta * C0.double()
        r = {0: r, 1: torch.relu(r), 2: torch.tanh(r)}[act]
        assert_close(C[:, :N], r, 2e-5, f"gemm_h16 act={act} beta={beta}")
        assert float(C[:, N:].abs().max() if ldc > N else 0.0) == 0.0      # TMA clips the N tail
    # 16-bit output (C = NULL, C16): rounded once after bias / activation; CTA-pair kernel only
    if M > 128 and N % 2 == 0:
        ld16 = (N + 7) // 8 * 8
        for act, c16_bf in ((0, False), (2, True)):
            dt = torch.bfloat16 if c16_bf else torch.float16
            C16 = torch.zeros(M, ld16, device=DEV, dtype=dt)
            use_bias = bias if N % 4 == 0 else None
            ops.gemm_h16(Ad, lda, ta, Bd, ldb, tb, None, 0, M, N, K, a_bf16=bool(a_bf), b_bf16=bool(b_bf),
                         bias=None if use_bias is None else cu(use_bias), act=act, C16=C16, ldc16=ld16, c16_bf16=c16_bf)
            r = ref + (0 if use_bias is None else use_bias.double())
            r = {0: r, 2: torch.tanh(r)}[act]
            want = r.float().to(dt)                      # one rounding of the exact result
            got = C16[:, :N].cpu()
            ulp = 2.0 ** (-7 if c16_bf else -10)
            # within one 16-bit ulp of the exact value (+ the fp32 accumulation error, relative to the largest entry)
            assert float(((got.double() - r).abs() - ulp * r.abs()).max()) < 1e-4 * float(ref.abs().max())
            if act == 0 and K <= 1024:   # (long fp32 accumulations move more results across a rounding boundary)
                assert float((got != want).float().mean()) < 0.02   # and almost always the correctly rounded value
            assert float(C16[:, N:].abs().max() if ld16 > N else 0.0) == 0.0
    # cast helper
    x = torch.randn(37, 84, generator=g)
    x16 = ops.cast16(cu(x))
    assert x16.shape == (37, 88) and float(x16[:, 84:].abs().max()) == 0.0
    assert torch.equal(x16[:, :84].cpu(), x.to(torch.float16))


@pytest.mark.gpu
def test_graph_replay_matches_eager_training():
    """CUDA-graph replay of the whole trainer step (fwd + loss + bwd + clip + AdamW) follows the eager step on changing
    data (dropout 0 so neither draws masks).  Float atomics in the gradient reductions make two runs differ in the
    last bits and early Adam steps (update ~ lr * sign(g)) amplify that, hence the loose per-step tolerance; the
    device-side step counter / learning rate are checked exactly in test_adamw_device_step_and_lr."""
    import copy
    import deer_b200
    from deer_b200.trainer import DEERDataParallelTrainer
    torch.manual_seed(3)
    base = deer_b200.CompleteDEERModel(deer_b200.ModelConfig(dropout=0.0)).cuda().train()
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    t1 = DEERDataParallelTrainer(m1, learning_rate=1e-3)
    t2 = DEERDataParallelTrainer(m2, learning_rate=1e-3)
    g = torch.Generator().manual_seed(0)

    def batch():
        return {"audio_features": torch.randn(32, 84, generator=g).cuda(),
                "video_features": torch.randn(32, 256, generator=g).cuda(),
                "text_features": torch.randn(32, 768, generator=g).cuda(),
                "targets": torch.tanh(torch.randn(32, 3, generator=g)).cuda()}

    data = [batch() for _ in range(7)]
    for i, b in enumerate(data):
        if i == 4:
            t1.lr = 5e-4
            t2.lr = 5e-4
        l1 = t1.train_step(b).clone()
        l2 = t2.train_step_auto(b).clone()       # eager, eager, then captured + replayed
        assert torch.allclose(l1, l2, rtol=5e-3, atol=1e-5), (i, l1, l2)
    assert t2._auto and next(iter(t2._auto.values()))["replay"] is not None
    assert int(t1.step_tensor) == int(t2.step_tensor) == 7
    p1, p2 = t1.flat.params, t2.flat.params
    assert float((p1 - p2).norm() / p1.norm()) <= 1e-3


@pytest.mark.gpu
def test_adamw_device_step_and_lr():
    """deer_adamw with the step count and learning rate read from device memory (graph replay) == host scalars."""
    from deer_b200._lib import call, ptr
    g = torch.Generator().manual_seed(1)
    n = 10007
    p0 = torch.randn(n, generator=g).cuda()
    gr = torch.randn(n, generator=g).cuda() * 0.1
    m0 = torch.randn(n, generator=g).cuda() * 0.01
    v0 = torch.rand(n, generator=g).cuda() * 0.01
    ss = (gr * gr).sum().reshape(1)
    for step in (1, 2, 57, 1000):
        pa, ma, va = p0.clone(), m0.clone(), v0.clone()
        pb, mb, vb = p0.clone(), m0.clone(), v0.clone()
        call("deer_adamw", ptr(pa), ptr(gr), ptr(ma), ptr(va), n, 3e-4 * 0.5, 0.9, 0.999, 1e-8, 1e-5, step, ptr(ss), 1.0,
             1.0, None, None)
        st = torch.tensor([step - 1], device="cuda", dtype=torch.int64)
        lr = torch.tensor([3e-4], device="cuda", dtype=torch.float32)
        call("deer_adamw", ptr(pb), ptr(gr), ptr(mb), ptr(vb), n, 0.5, 0.9, 0.999, 1e-8, 1e-5, 0, ptr(ss), 1.0, 1.0,
             st.data_ptr(), ptr(lr))
        assert torch.equal(ma, mb) and torch.equal(va, vb)
        assert float((pa - pb).abs().max()) <= 2.5e-7 * float(p0.abs().max()), step   # 1-2 ulp of the parameter
        # and against torch.optim.AdamW semantics (training.py:121-150) in fp64
        clip = min(1.0, 1.0 / (float(ss.sqrt()) + 1e-6))
        gd, pd_, md, vd = gr.double() * clip, p0.double(), m0.double(), v0.double()
        md = 0.9 * md + 0.1 * gd
        vd = 0.999 * vd + 0.001 * gd * gd
        lr_ = 1.5e-4
        want = pd_ * (1 - lr_ * 1e-5) - lr_ / (1 - 0.9 ** step) * md / (vd.sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8)
        assert float((pb.double() - want).abs().max()) <= 1e-4 * lr_ + 2.5e-7 * float(p0.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("M,N,K,ta,tb", [(3000, 256, 512, 0, 1), (1300, 520, 200, 0, 0), (1024, 512, 9000, 1, 0),
                                         (700, 300, 132, 1, 1), (12800, 512, 1536, 0, 1), (300, 128, 64, 0, 1)])
def test_gemm_tf32_engines(M, N, K, ta, tb, pair):
    """TF32 tcgen05 engines through deer_gemm: the 128x128-tile kernel and the cta_group::2 CTA-pair kernel (256x256
    tiles), all operand majors, M/N/K tails, bias + activation and the accumulating (beta = 1, split-K) epilogue."""
    from deer_b200 import _lib
    _lib.set_option(6, pair)
    try:
        g = torch.Generator().manual_seed(M + N + K)
        lda = ((M if ta else K) + 3) // 4 * 4
        ldb = ((K if tb else N) + 3) // 4 * 4
        A = torch.zeros((K if ta else M), lda)
        sc = 1.5 * K ** -0.25      # pre-activations of order 1 (a saturated tanh would hide / amplify nothing useful)
        A[:, :(M if ta else K)] = torch.randn((K if ta else M), (M if ta else K), generator=g) * sc
        Bm = torch.zeros((N if tb else K), ldb)
        Bm[:, :(K if tb else N)] = torch.randn((N if tb else K), (K if tb else N), generator=g) * sc
        opA = (A[:, :M].t() if ta else A[:, :K]).double()
        opB = (Bm[:, :K].t() if tb else Bm[:, :N]).double()
        ref = opA @ opB
        bias = torch.randn(N, generator=g)
        C0 = torch.randn(M, N, generator=g)
        ldc = (N + 3) // 4 * 4
        for act, beta in ((0, 0.0), (2, 0.0), (0, 1.0)):
            C = torch.zeros(M, ldc, device=DEV)
            C[:, :N] = cu(C0)
            use_bias = bias if N % 4 == 0 else None
            ops.gemm(cu(A), lda, ta, cu(Bm), ldb, tb, C, ldc, M, N, K, bias=None if use_bias is None else cu(use_bias),
                     act=act, beta=beta, engine=ops.ENGINE_TF32)
            r = ref + (0 if use_bias is None else use_bias.double()) + beta * C0.double()
            r = torch.tanh(r) if act == 2 else r
            assert_close(C[:, :N], r, 1e-3, f"tf32 pair={pair} act={act} beta={beta}")
    finally:
        _lib.set_option(6, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,Cin,Cout", [(32, 12, 64, 512), (300, 50, 512, 512), (7, 50, 128, 260)])
def test_conv1d_sliding_window_matches_im2col_and_torch(B, T, Cin, Cout):
    """Conv1d(k=3, padding=1) through overlapping-row TMA windows of a zero-row-padded copy (no im2col matrix): same
    result as the im2col + GEMM path and as torch's fp64 conv1d (encoders.py:450-459), forward and all gradients."""
    from deer_b200 import _lib
    g = torch.Generator().manual_seed(B * T + Cin)
    x = torch.randn(B, T, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, generator=g) * (3 * Cin) ** -0.5
    b = torch.randn(Cout, generator=g) * 0.1
    pr = torch.randn(B, T, Cout, generator=g)
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.conv1d(xr.transpose(1, 2), wr, br, padding=1).transpose(1, 2)
    (yr * pr.double()).sum().backward()
    outs = []
    for window in (True, False):
        ops.set_conv_window(window)
        try:
            xc, wc, bc = (cu(t).requires_grad_(True) for t in (x, w, b))
            before = _lib.launch_count()
            y = ops.conv1d_k3(xc, wc, bc)
            (y * cu(pr)).sum().backward()
            outs.append((y.detach(), xc.grad, wc.grad, bc.grad, _lib.launch_count() - before))
        finally:
            ops.set_conv_window(True)
    for got in outs:
        assert_close(got[0], yr, 1e-3, "conv y")
        assert_close(got[1], xr.grad, 1e-3, "conv dx")
        assert_close(got[2], wr.grad, 1e-3, "conv dw")
        assert_close(got[3], br.grad, 1e-3, "conv db")
    for a, c in zip(outs[0][:4], outs[1][:4]):
        assert_close(a, c, 1e-3, "window vs im2col")


def test_stream_concurrency_switches_do_not_change_results():
    """The audio encoder on its own high-priority stream, the deferred small-layer weight gradients on a third stream
    and the FP16 hand-off between LSTM layers only reorder / overlap launches: gradients and losses are identical to
    the single-stream schedule (same kernels, same operands, disjoint accumulation targets)."""
    from deer_b200.trainer import DEERDataParallelTrainer
    from gen_common import seq_inputs

    res = []
    for branch, defer, text in ((True, True, True), (False, False, False), (True, False, True), (False, True, True),
                                (True, True, False)):
        ops.set_branch_streams(branch)
        ops.set_defer_wgrad(defer)
        ops.set_text_stream(text)        # the text encoder on a third stream beside the video encoder
        try:
            torch.manual_seed(3)
            model = deer_b200.SequenceDEERModel(dropout=0.0).to(DEV).train()
            tr = DEERDataParallelTrainer(model)
            b = [t.float().to(DEV) for t in seq_inputs(6, 9, 5, 7, seed=11)]
            batch = {"audio_features": b[0], "video_features": b[1], "text_features": b[2], "attention_mask": b[3],
                     "linguistic_features": b[4], "targets": b[5]}
            for _ in range(2):          # twice: stream reuse across steps
                tr.flat.grads.zero_()
                losses = tr.forward_backward(batch)
            torch.cuda.synchronize()
            res.append((tr.flat.grads.clone(), losses.clone()))
        finally:
            ops.set_branch_streams(True)
            ops.set_defer_wgrad(True)
            ops.set_text_stream(True)
    assert float(res[0][0].norm()) > 0
    for g, l in res[1:]:      # not bit-wise: several gradients are accumulated with floating-point atomics
        assert_close(g, res[0][0], 1e-5, "flat gradients")
        assert_close(l, res[0][1], 1e-6, "losses")


def test_lstm_fp16_handoff_between_layers_is_exact():
    """Inference / no-dropout: layer 0's recurrence kernel writes the FP16 copy of h that layer 1's input projection
    consumes (no cast pass).  Same rounding (RN to FP16 of the same fp32 h) -> bit-identical encoder output."""
    torch.manual_seed(1)
    enc = deer_b200.EnhancedAudioEncoder({"dropout": 0.3}).to(DEV).eval()
    x = torch.randn(40, 30, 84, device=DEV)
    with torch.no_grad():
        h_fast = enc.lstm_forward(x)
        h = ops.to_time_major(x)
        for l in range(enc.num_layers):           # the same two layers, each casting its own input
            h = ops.bilstm_layer(h, *enc._layer_weights(l))
    assert torch.equal(h_fast, h)
    # training without dropout takes the hand-off too and must give the same gradients as the cast path
    enc.train()
    enc.dropout = 0.0
    grads = []
    for handoff in (True, False):
        for p in enc.parameters():
            p.grad = None
        xi = x.clone().requires_grad_(True)
        if handoff:
            h = enc.lstm_forward(xi)
        else:
            h = ops.to_time_major(xi)
            for l in range(enc.num_layers):
                h = ops.bilstm_layer(h, *enc._layer_weights(l))
        (h * torch.linspace(-1, 1, h.numel(), device=DEV).view_as(h)).sum().backward()
        grads.append([xi.grad.clone()] + [p.grad.clone() for p in enc.lstm.parameters()])
    for a, b in zip(*grads):      # atomically accumulated bias gradients: equal up to summation order
        assert_close(a, b, 1e-5, "lstm gradients")


@pytest.mark.parametrize("B,D", [(300001, 3), (70001, 1), (1000, 2)])
def test_fused_loss_pipelined_and_plain_trips_agree(B, D):
    """Several grid-stride trips per thread (software-pipelined loads, DEER_OPT_NIG_PIPELINE, default) against the plain
    load -> compute trips and against the fp64 oracle on the loss value: element pairs, odd tails, D = 1 / 2 / 3."""
    from deer_b200 import _lib
    g = torch.Generator().manual_seed(B + D)
    e = torch.randn(B, D, 4, generator=g, dtype=torch.float64) * 2.0
    e[:, :, 2] = e[:, :, 2].clamp(min=-5.0)
    y = torch.tanh(torch.randn(B, D, generator=g, dtype=torch.float64))
    res = []
    try:
        for pipe in (1, 0):
            _lib.set_option(8, pipe)
            losses, dE, nig, _ = ops.nig_loss_raw(cu(e), None, cu(y), want_nig=True, want_grad=True)
            torch.cuda.synchronize()
            res.append((losses.clone(), dE.clone(), nig.clone()))
    finally:
        _lib.set_option(8, 1)
    (l1, d1, n1), (l0, d0, n0) = res
    assert_close(l1, l0, 1e-5, "losses")          # statistics are accumulated with atomics: equal up to summation order
    assert_close(d1, d0, 1e-5, "gradient")
    assert torch.equal(n1, n0)                    # the head outputs are per-element: bit-identical
    if D == 3:
        pred = {}
        for i, dname in enumerate(("valence", "arousal", "dominance")):
            for k, v in O.nig_from_evidence(e[:, i:i + 1, :]).items():
                pred[f"{dname}_{k}"] = v
        ref = float(O.multitask_deer_loss(pred, y)["total_loss"])
        assert abs(float(l1[-1]) - ref) <= TOL * abs(ref)


@pytest.mark.parametrize("time_major,masked", [(True, False), (False, False), (False, True)])
def test_fused_scorer_pool_matches_separate_nodes(time_major, masked):
    """ops.scorer_pool (one autograd node; the scorer's dx GEMM accumulates onto the pooling kernel's dx) against the
    separate Linear-tanh / rowdot / attn_pool nodes whose input gradients autograd adds: same output and gradients."""
    g = torch.Generator().manual_seed(5)
    B, T, D, Hd = 37, 21, 64, 32
    x0 = torch.randn((T, B, D) if time_major else (B, T, D), generator=g)
    w1, b1 = torch.randn(Hd, D, generator=g) * 0.2, torch.randn(Hd, generator=g) * 0.1
    w2, b2 = torch.randn(1, Hd, generator=g) * 0.3, torch.randn(1, generator=g)
    mask = (torch.rand(B, T, generator=g) > 0.3).float() if masked else None
    if masked:
        mask[:, 0] = 1.0
    pr = torch.randn(B, D, generator=g)
    res = []
    for fused in (True, False):
        xs = cu(x0).requires_grad_(True)
        ps = [cu(t).requires_grad_(True) for t in (w1, b1, w2, b2)]
        m = None if mask is None else cu(mask)
        if fused:
            out, wts = ops.scorer_pool(xs, *ps, m, time_major)
        else:
            hidden = ops.linear(xs, ps[0], ps[1], "tanh")
            s = ops.rowdot(hidden, ps[2].view(-1), ps[3])
            out, wts = ops.attn_pool(xs.permute(1, 0, 2), s.permute(1, 0), m) if time_major else ops.attn_pool(xs, s, m)
        (out * cu(pr)).sum().backward()
        res.append([out.detach(), wts.detach(), xs.grad] + [p.grad for p in ps])
    for i, (a, b) in enumerate(zip(*res)):
        if i == len(res[0]) - 1:
            # d/d b2 is sum_t ds = 0 in exact arithmetic (softmax is shift invariant): pure rounding noise in both
            assert float(a.abs().max()) < 1e-4 and float(b.abs().max()) < 1e-4
            continue
        # bias gradients are atomically accumulated: equal up to the summation order
        assert_close(a, b, 1e-4, "fused vs separate")
    # and against torch fp64
    xd = x0.double().requires_grad_(True)
    xb = xd.permute(1, 0, 2) if time_major else xd
    sc = torch.tanh(xb @ w1.double().t() + b1.double()) @ w2.double().view(-1) + b2.double()
    p = torch.softmax(sc, dim=1)
    if masked:
        p = p * mask.double()
        p = p / (p.sum(dim=1, keepdim=True) + 1e-10)
    ref = (p.unsqueeze(-1) * xb).sum(dim=1)
    (ref * pr.double()).sum().backward()
    assert_close(res[0][0], ref, 1e-4, "pooled vs fp64")
    assert_close(res[0][2], xd.grad, 1e-3, "dx vs fp64")


@pytest.mark.parametrize("shared_input", [True, False])
@pytest.mark.parametrize("act", ["relu", "none"])
def test_grouped_linear_batched_matches_per_group_launches(shared_input, act):
    """Per-head grouped Linear as ONE batched GEMM launch (evenly spaced operands) against one launch per group:
    outputs, input gradients and weight / bias gradients."""
    g = torch.Generator().manual_seed(9)
    G, M, K, N = 3, 50, 48, 32
    wbuf = torch.randn(G, N * K + N + 16, generator=g) * 0.2          # [W | b | pad] per head, evenly spaced
    xin = torch.randn(M, K, generator=g) if shared_input else torch.randn(M, G, K, generator=g)
    pr = torch.randn(M, G, N, generator=g)
    res = []
    for batched in (True, False):
        ops.set_grouped_batched(batched)
        try:
            wb = cu(wbuf).requires_grad_(True)
            ws = [wb[i, :N * K].view(N, K) for i in range(G)]
            bs = [wb[i, N * K:N * K + N] for i in range(G)]
            x = cu(xin).requires_grad_(True)
            xs = [x] * G if shared_input else [x[:, i] for i in range(G)]
            out = ops.grouped_linear(xs, ws, bs, act)
            (out * cu(pr)).sum().backward()
            res.append((out.detach(), x.grad, wb.grad))
        finally:
            ops.set_grouped_batched(True)
    for a, b in zip(*res):
        assert_close(a, b, 1e-5, "batched vs per-group")
