"""Pin the CPU oracle (oracle/deer_oracle.py) against golden vectors produced by
the unmodified reference modules (tests/golden/make_golden.py).  float64 on both
sides, so agreement is to round-off: rtol 1e-9."""
import numpy as np
import pytest
import torch

from gen_common import det_normal, grad_summary, nig_inputs, pooled_inputs, probe, seq_inputs
from oracle import deer_oracle as O

RT, AT = 1e-9, 1e-11


def close(a, b, rt=RT, at=AT):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rt, atol=at), float((a - b).abs().max())


def leaf_sd(fx):
    return {k: (v.clone().requires_grad_(True) if v.dtype == torch.float64 else v) for k, v in fx.state_dict().items()}


def check_grads(fx, sd, seed, prefix=""):
    n = 0
    for k in fx.keys("grad:"):
        name = k[5:]
        g = sd[prefix + name].grad
        g = torch.zeros_like(sd[prefix + name]) if g is None else g
        close(g, fx.t(k), 1e-8, 1e-10)
        n += 1
    for k in fx.keys("gsum:"):
        name = k[5:]
        g = sd[prefix + name].grad
        g = torch.zeros_like(sd[prefix + name]) if g is None else g
        nrm, dot = grad_summary(g, name, seed)
        ref = fx.arrays[k]
        assert abs(nrm - ref[0]) <= 1e-8 * max(1, abs(ref[0])), (name, nrm, ref)
        assert abs(dot - ref[1]) <= 1e-8 * max(1, abs(ref[0])), (name, dot, ref)
        n += 1
    assert n > 0


@pytest.mark.parametrize("name", ["audio_small", "audio_full_t12"])
def test_audio_encoder(golden, name):
    fx = golden(name)
    m = fx.meta
    sd = leaf_sd(fx)
    x = seq_inputs(m["B"], m["T"], 2, 2, seed=m["seed"])[0].clone().requires_grad_(True)
    close(O.bilstm(x, sd), fx.t("lstm_out"))
    y = O.audio_encoder(x, sd)
    close(y, fx.t("out"))
    (y * probe("audio_out", y.shape, m["seed"])).sum().backward()
    close(x.grad, fx.t("dx"), 1e-8, 1e-10)
    check_grads(fx, sd, m["seed"])


@pytest.mark.parametrize("name", ["video_small_train", "video_small_eval", "video_small_f1"])
def test_video_encoder(golden, name):
    fx = golden(name)
    m = fx.meta
    sd = leaf_sd(fx)
    x = seq_inputs(m["B"], 2, m["F"], 2, Dv=m["din"], seed=m["seed"])[1].clone().requires_grad_(True)
    y = O.video_encoder(x, sd, training=m["training"])
    close(y, fx.t("out"))
    (y * probe("video_out", y.shape, m["seed"])).sum().backward()
    close(x.grad, fx.t("dx"), 1e-8, 1e-10)
    check_grads(fx, sd, m["seed"])


def test_text_encoder(golden):
    fx = golden("text_small")
    m = fx.meta
    sd = leaf_sd(fx)
    _, _, tok, mask, ling, _ = seq_inputs(m["B"], 2, 2, m["T"], seed=m["seed"])
    tok = tok.clone().requires_grad_(True)
    ling = ling.clone().requires_grad_(True)
    y = O.text_encoder(tok, mask, ling, sd)
    close(y, fx.t("out"))
    (y * probe("text_out", y.shape, m["seed"])).sum().backward()
    close(tok.grad, fx.t("dtok"), 1e-8, 1e-10)
    close(ling.grad, fx.t("dling"), 1e-8, 1e-10)
    check_grads(fx, sd, m["seed"])


def test_fusion(golden):
    fx = golden("fusion_small")
    da, dv, dt, fd, idim, heads = fx.meta["dims"]
    B, seed = fx.meta["B"], fx.meta["seed"]
    sd = leaf_sd(fx)
    a = det_normal("in:fa", (B, da), seed).requires_grad_(True)
    v = det_normal("in:fv", (B, dv), seed).requires_grad_(True)
    t = det_normal("in:ft", (B, dt), seed).requires_grad_(True)
    out = O.hierarchical_fusion(a, v, t, sd, heads=heads)
    keys = ("fused_features", "audiovisual_features", "trimodal_features", "trimodal_attention_weights")
    for k in keys:
        close(out[k], fx.t(k))
    close(out["av_attention_weights"]["audio_to_video"], fx.t("a2v"))
    close(out["av_attention_weights"]["video_to_audio"], fx.t("v2a"))
    sum((out[k] * probe(k, out[k].shape, seed)).sum() for k in keys).backward()
    close(a.grad, fx.t("da"), 1e-8, 1e-10)
    close(v.grad, fx.t("dv"), 1e-8, 1e-10)
    close(t.grad, fx.t("dt"), 1e-8, 1e-10)
    check_grads(fx, sd, seed)
    # parameters the reference never touches (uncertainty_gate.*) have no grad
    for k in fx.keys("hasgrad:"):
        if not bool(fx.arrays[k]):
            assert sd[k[8:]].grad is None


def test_head(golden):
    fx = golden("head_small")
    m = fx.meta
    sd = leaf_sd(fx)
    x = det_normal("in:hx", (m["B"], m["din"]), m["seed"]).requires_grad_(True)
    out = O.multidim_deer(x, sd)
    for k, v in out.items():
        close(v, fx.t(k))
    sum((v * probe(k, v.shape, m["seed"])).sum() for k, v in out.items()).backward()
    close(x.grad, fx.t("dx"), 1e-8, 1e-10)
    check_grads(fx, sd, m["seed"])


@pytest.mark.parametrize("name", ["loss_b64", "loss_b1000", "loss_b3"])
def test_losses(golden, name):
    fx = golden(name)
    e, y = nig_inputs(fx.meta["B"], fx.meta["seed"])
    e = e.clone().requires_grad_(True)
    p = O.nig_from_evidence(e)
    pred = {}
    for i, d in enumerate(O.DIMS):
        for k in ("mu", "nu", "alpha", "beta"):
            pred[f"{d}_{k}"] = p[k][:, i:i + 1]
    out = O.multitask_deer_loss(pred, y)
    for k in fx.keys("mt:"):
        if k == "mt:devidence":
            continue
        close(out[k[3:]], fx.t(k).reshape(()), 1e-10, 1e-12)
    out["total_loss"].backward(retain_graph=True)
    close(e.grad, fx.t("mt:devidence"), 1e-8, 1e-12)
    single = O.deer_loss(p["mu"][:, 0:1], p["nu"][:, 0:1], p["alpha"][:, 0:1], p["beta"][:, 0:1], y[:, 0:1])
    for k in fx.keys("l1:"):
        close(single[k[3:]], fx.t(k).reshape(()), 1e-10, 1e-12)
    e.grad = None
    am = O.amini_deer_loss(p["mu"], p["nu"], p["alpha"], p["beta"], y)
    for k in fx.keys("l3:"):
        if k != "l3:devidence":
            close(am[k[3:]], fx.t(k).reshape(()), 1e-10, 1e-12)
    am["total_loss"].backward()
    close(e.grad, fx.t("l3:devidence"), 1e-8, 1e-12)


def test_sequence_composite(golden):
    fx = golden("seq_full_b4")
    m = fx.meta
    sd = leaf_sd(fx)
    audio, video, text, mask, ling, y = seq_inputs(m["B"], m["Ta"], m["Tv"], m["Tt"], seed=m["seed"])
    out, loss = O.sequence_model_loss(audio, video, text, mask, ling, y, sd, training=True)
    for k in ("audio_encoded", "video_encoded", "text_encoded", "fused_features"):
        close(out[k], fx.t(k), 1e-8, 1e-10)
    for k in fx.keys("out:"):
        close(out[k[4:]], fx.t(k), 1e-8, 1e-10)
    for k in fx.keys("loss:"):
        close(loss[k[5:]], fx.t(k).reshape(()), 1e-8, 1e-10)
    loss["total_loss"].backward()
    check_grads(fx, sd, m["seed"])
    for k, has in m["has_grad"].items():
        if not has:
            assert sd[k].grad is None, k


def test_pooled_model(golden):
    fx = golden("pooled_b16")
    m = fx.meta
    sd = leaf_sd(fx)
    a, v, t, y = pooled_inputs(m["B"], m["seed"])
    out = O.pooled_model(a, v, t, sd)
    for k in fx.keys("out:"):
        close(out[k[4:]], fx.t(k), 1e-8, 1e-10)
    loss = O.multitask_deer_loss(O.pooled_loss_inputs(out), y)
    for k in fx.keys("loss:"):
        close(loss[k[5:]], fx.t(k).reshape(()), 1e-8, 1e-10)
    loss["total_loss"].backward()
    check_grads(fx, sd, m["seed"])
    for k, has in m["has_grad"].items():
        if not has or float(fx.arrays.get("gsum:" + k, np.ones(2))[0]) == 0.0:
            g = sd[k].grad
            assert g is None or float(g.abs().max()) == 0.0, k


def test_linguistic_features(golden):
    fx = golden("ling")
    ids = torch.from_numpy(fx.arrays["ids"])
    mask = torch.from_numpy(fx.arrays["mask"])
    got = O.linguistic_features(ids, mask)
    assert torch.equal(got, torch.from_numpy(fx.arrays["feats"]))


def test_torch_baseline_matches_oracle(golden):
    """The CPU baseline that bench.py times (oracle/torch_baseline.py, stock torch.nn layers as the reference uses)
    computes the same numbers as the golden-pinned oracle."""
    from oracle import torch_baseline as TB
    fx = golden("seq_full_b4")
    m = fx.meta
    model = TB.SequenceBaseline(dropout=0.0).double()
    missing, unexpected = model.load_state_dict(fx.state_dict(), strict=False)
    assert not missing and all(k.startswith("fusion.uncertainty_gate") for k in unexpected), (missing, unexpected)
    model.train()
    audio, video, text, mask, ling, y = seq_inputs(m["B"], 40, 12, 16, seed=3)
    pred = model(audio, video, text, mask, ling)
    loss = TB.multitask_loss(pred, y)
    ref_out, ref_loss = O.sequence_model_loss(audio, video, text, mask, ling, y, fx.state_dict(), training=True)
    for d in O.DIMS:
        for k in ("mu", "nu", "alpha", "beta"):
            close(pred[f"{d}_{k}"], ref_out[f"{d}_{k}"], 1e-8, 1e-10)
    close(loss, ref_loss["total_loss"], 1e-8, 1e-10)
