"""Generate golden vectors by executing the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Nothing in tests/, smoke() or bench.py reads
/root/reference at run time; they read these fixtures.

Import shims follow SURVEY.md section 8c: a stub `librosa` module (absent here),
TRANSFORMERS_AVAILABLE=False so no BERT download, `enc.bert` replaced by a lambda
returning the supplied 768-D features, `extract_spatial_features` -> identity and
`spatial_projection[0]` swapped to the requested frame-feature width.
All reference modules run in float64 (`module.double()`), dropout 0.
"""
from __future__ import annotations

import importlib.machinery
import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_common import (det_state_dict, grad_summary, nig_inputs, pooled_inputs, probe,  # noqa: E402
                        seq_inputs)

REF = "/root/reference"


def import_reference():
    import transformers  # noqa: F401  (must be imported before the librosa stub)
    sys.path[:0] = [f"{REF}/src/models", f"{REF}/src/utils"]
    lib = types.ModuleType("librosa")
    lib.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
    sys.modules.setdefault("librosa", lib)
    import complete_project
    import deer
    import encoders
    import fusion
    import losses
    encoders.TRANSFORMERS_AVAILABLE = False
    return SimpleNamespace(encoders=encoders, fusion=fusion, deer=deer, losses=losses,
                           complete_project=complete_project)


def load_det(mod: nn.Module, seed: int, skip=()):
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items() if not k.startswith(tuple(skip))}
    sd = det_state_dict(shapes, seed)
    mod.load_state_dict(sd, strict=False)
    return shapes


def save(name, shapes, arrays, meta):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    arrays["__shapes__"] = np.frombuffer(json.dumps({k: list(v) for k, v in shapes.items()}).encode(), dtype=np.uint8)
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items() if not k.startswith("__")} if len(arrays) < 12 else len(arrays))


def put_grad(arr, k, g, seed, full_limit=8192):
    """Small gradients are stored whole; big ones as (norm, probe-dot) + first 64 entries."""
    if g.numel() <= full_limit:
        arr["grad:" + k] = g
    else:
        arr["gsum:" + k] = np.array(grad_summary(g, k, seed))
        arr["ghead:" + k] = g.flatten()[:64]


def grads_of(mod, skip=()):
    return {k: p.grad for k, p in mod.named_parameters() if p.grad is not None and not k.startswith(tuple(skip))}


def make_audio(R, hidden, B, T, seed, name, full_grads):
    enc = R.encoders.EnhancedAudioEncoder({"hidden_dim": hidden, "dropout": 0.0}).double()
    shapes = load_det(enc, seed)
    x = seq_inputs(B, T, 2, 2, seed=seed)[0].clone().requires_grad_(True)
    lstm_out, _ = enc.lstm(x)
    y = enc(x)
    (y * probe("audio_out", y.shape, seed)).sum().backward()
    arr = {"out": y, "lstm_out": lstm_out, "dx": x.grad}
    for k, g in grads_of(enc).items():
        put_grad(arr, k, g, seed, 8192 if full_grads else 0)
    save(name, shapes, arr, {"hidden": hidden, "B": B, "T": T, "seed": seed})


def make_video(R, hidden, din, B, F, seed, name, training):
    enc = R.encoders.EnhancedVideoEncoder({"hidden_dim": hidden, "dropout": 0.0})
    enc.extract_spatial_features = lambda v: v
    enc.spatial_projection[0] = nn.Linear(din, hidden)
    enc = enc.double()
    enc.train(training)
    shapes = load_det(enc, seed, skip=("spatial_backbone",))
    x = seq_inputs(B, 2, F, 2, Dv=din, seed=seed)[1].clone().requires_grad_(True)
    y = enc(x)
    (y * probe("video_out", y.shape, seed)).sum().backward()
    arr = {"out": y, "dx": x.grad}
    for k, g in grads_of(enc, skip=("spatial_backbone",)).items():
        put_grad(arr, k, g, seed)
    if training:
        for k, v in enc.state_dict().items():
            if "running" in k and not k.startswith("spatial_backbone"):
                arr["post:" + k] = v
    save(name, shapes, arr, {"hidden": hidden, "din": din, "B": B, "F": F, "seed": seed, "training": training})


def make_text(R, hidden, B, T, seed, name):
    enc = R.encoders.EnhancedTextEncoder({"hidden_dim": hidden, "dropout": 0.0}).double()
    skip = ("embedding", "positional_encoding")
    shapes = load_det(enc, seed, skip=skip)
    _, _, tok, mask, ling, _ = seq_inputs(B, 2, 2, T, seed=seed)
    tok = tok.clone().requires_grad_(True)
    ling = ling.clone().requires_grad_(True)
    enc.bert = lambda input_ids, attention_mask: SimpleNamespace(last_hidden_state=tok)
    enc.extract_linguistic_features = lambda ids, m: ling
    ids = torch.zeros(B, T, dtype=torch.long)
    y = enc(ids, mask)
    (y * probe("text_out", y.shape, seed)).sum().backward()
    arr = {"out": y, "dtok": tok.grad, "dling": ling.grad}
    for k, g in grads_of(enc, skip=skip).items():
        put_grad(arr, k, g, seed)
    save(name, shapes, arr, {"hidden": hidden, "B": B, "T": T, "seed": seed})


def make_fusion(R, da, dv, dt, fd, idim, heads, B, seed, name):
    fus = R.fusion.HierarchicalMultimodalFusion(da, dv, dt, fusion_dim=fd, intermediate_dim=idim,
                                                num_attention_heads=heads, dropout=0.0).double()
    shapes = load_det(fus, seed)
    from gen_common import det_normal
    a = det_normal("in:fa", (B, da), seed).requires_grad_(True)
    v = det_normal("in:fv", (B, dv), seed).requires_grad_(True)
    t = det_normal("in:ft", (B, dt), seed).requires_grad_(True)
    out = fus(a, v, t)
    s = sum((out[k] * probe(k, out[k].shape, seed)).sum() for k in
            ("fused_features", "audiovisual_features", "trimodal_features", "trimodal_attention_weights"))
    s.backward()
    arr = {"fused_features": out["fused_features"], "audiovisual_features": out["audiovisual_features"],
           "trimodal_features": out["trimodal_features"],
           "trimodal_attention_weights": out["trimodal_attention_weights"],
           "a2v": out["av_attention_weights"]["audio_to_video"], "v2a": out["av_attention_weights"]["video_to_audio"],
           "da": a.grad, "dv": v.grad, "dt": t.grad}
    for k, p in fus.named_parameters():
        put_grad(arr, k, p.grad if p.grad is not None else torch.zeros_like(p), seed)
        arr["hasgrad:" + k] = np.array(p.grad is not None)
    save(name, shapes, arr, {"dims": [da, dv, dt, fd, idim, heads], "B": B, "seed": seed})


def make_head(R, din, hidden, B, seed, name):
    head = R.deer.MultiDimensionalDEER(din, 3, hidden, dropout=0.0).double()
    shapes = load_det(head, seed)
    from gen_common import det_normal
    x = det_normal("in:hx", (B, din), seed).requires_grad_(True)
    out = head(x)
    s = sum((v * probe(k, v.shape, seed)).sum() for k, v in out.items())
    s.backward()
    arr = dict(out)
    arr["dx"] = x.grad
    for k, g in grads_of(head).items():
        arr["grad:" + k] = g
    save(name, shapes, arr, {"din": din, "hidden": hidden, "B": B, "seed": seed})


def make_loss(R, B, seed, name):
    e, y = nig_inputs(B, seed)
    e = e.clone().requires_grad_(True)
    mu = e[..., 0]
    nu = torch.nn.functional.softplus(e[..., 1]) + 1e-6
    al = torch.nn.functional.softplus(e[..., 2]) + 1.0
    be = torch.nn.functional.softplus(e[..., 3]) + 1e-6
    pred = {}
    for i, d in enumerate(("valence", "arousal", "dominance")):
        pred[f"{d}_mu"], pred[f"{d}_nu"] = mu[:, i:i + 1], nu[:, i:i + 1]
        pred[f"{d}_alpha"], pred[f"{d}_beta"] = al[:, i:i + 1], be[:, i:i + 1]
    mt = R.losses.MultiTaskDEERLoss()
    out = mt(pred, y)
    out["total_loss"].backward(retain_graph=True)
    arr = {"mt:" + k: (v if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in out.items()}
    arr["mt:devidence"] = e.grad.clone()
    # single-dimension losses.DEERLoss and deer.DEERLoss (L1 / L3) on dimension 0
    e.grad = None
    single = R.losses.DEERLoss()({"mu": mu[:, 0:1], "nu": nu[:, 0:1], "alpha": al[:, 0:1], "beta": be[:, 0:1]}, y[:, 0:1])
    for k, v in single.items():
        arr["l1:" + k] = v if torch.is_tensor(v) else torch.tensor(float(v))
    am = R.deer.DEERLoss()({"mu": mu, "nu": nu, "alpha": al, "beta": be}, y)
    am["total_loss"].backward()
    for k, v in am.items():
        arr["l3:" + k] = v
    arr["l3:devidence"] = e.grad.clone()
    # the same multitask loss in float32, to record what "reference fp32" deviates by
    e32 = e.detach().float()
    p32 = {}
    for i, d in enumerate(("valence", "arousal", "dominance")):
        p32[f"{d}_mu"] = e32[:, i:i + 1, 0]
        p32[f"{d}_nu"] = torch.nn.functional.softplus(e32[:, i:i + 1, 1]) + 1e-6
        p32[f"{d}_alpha"] = torch.nn.functional.softplus(e32[:, i:i + 1, 2]) + 1.0
        p32[f"{d}_beta"] = torch.nn.functional.softplus(e32[:, i:i + 1, 3]) + 1e-6
    arr["mt32:total_loss"] = mt(p32, y.float())["total_loss"]
    save(name, {}, arr, {"B": B, "seed": seed})


def build_seq_reference(R, seed):
    a = R.encoders.EnhancedAudioEncoder({"hidden_dim": 512, "dropout": 0.0})
    v = R.encoders.EnhancedVideoEncoder({"hidden_dim": 512, "dropout": 0.0})
    v.extract_spatial_features = lambda z: z
    v.spatial_projection[0] = nn.Linear(256, 512)
    t = R.encoders.EnhancedTextEncoder({"hidden_dim": 512, "dropout": 0.0})
    f = R.fusion.HierarchicalMultimodalFusion(512, 512, 512, dropout=0.0)
    h = R.deer.MultiDimensionalDEER(512, 3, 256, dropout=0.0)
    mods = {"audio_encoder": a.double(), "video_encoder": v.double(), "text_encoder": t.double(),
            "fusion": f.double(), "deer": h.double()}
    skip = ("spatial_backbone", "embedding", "positional_encoding")
    shapes = {}
    for name, m in mods.items():
        sub = {k: tuple(p.shape) for k, p in m.state_dict().items() if not k.startswith(skip)}
        sd = det_state_dict({f"{name}.{k}": s for k, s in sub.items()}, seed)
        m.load_state_dict({k[len(name) + 1:]: val for k, val in sd.items()}, strict=False)
        shapes.update({f"{name}.{k}": s for k, s in sub.items()})
    return mods, shapes, skip


def make_seq_full(R, B, Ta, Tv, Tt, seed, name):
    mods, shapes, skip = build_seq_reference(R, seed)
    audio, video, text, mask, ling, y = seq_inputs(B, Ta, Tv, Tt, seed=seed)
    mods["text_encoder"].bert = lambda input_ids, attention_mask: SimpleNamespace(last_hidden_state=text)
    mods["text_encoder"].extract_linguistic_features = lambda ids, m: ling
    for m in mods.values():
        m.train()
    ae = mods["audio_encoder"](audio)
    ve = mods["video_encoder"](video)
    te = mods["text_encoder"](torch.zeros(B, Tt, dtype=torch.long), mask)
    fo = mods["fusion"](ae, ve, te)
    out = mods["deer"](fo["fused_features"])
    loss = R.losses.MultiTaskDEERLoss()(out, y)
    loss["total_loss"].backward()
    arr = {"audio_encoded": ae, "video_encoded": ve, "text_encoded": te, "fused_features": fo["fused_features"]}
    arr.update({"out:" + k: v for k, v in out.items()})
    arr.update({"loss:" + k: (v if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in loss.items()})
    nz = {}
    for mname, m in mods.items():
        for k, p in m.named_parameters():
            if k.startswith(skip):
                continue
            full = f"{mname}.{k}"
            if p.grad is None:
                nz[full] = False
                continue
            nz[full] = True
            arr["gsum:" + full] = np.array(grad_summary(p.grad, full, seed))
    save(name, shapes, arr, {"B": B, "Ta": Ta, "Tv": Tv, "Tt": Tt, "seed": seed, "has_grad": nz})


def make_pooled(R, B, seed, name):
    cp = R.complete_project
    model = cp.CompleteDEERModel(cp.ModelConfig(dropout=0.0)).double()
    model.train()
    # attention dropout inside MultiHeadAttention is config.dropout -> 0.0; UncertaintyEstimator has a hard-coded
    # Dropout(0.2) (complete_project.py:193) which must be disabled for a deterministic golden
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    shapes = load_det(model, seed)
    a, v, t, y = pooled_inputs(B, seed)
    out = model(a, v, t)
    loss = R.losses.MultiTaskDEERLoss()(out, y)
    loss["total_loss"].backward()
    arr = {"out:" + k: val for k, val in out.items()}
    arr.update({"loss:" + k: (val if torch.is_tensor(val) else torch.tensor(float(val))) for k, val in loss.items()})
    nz = {}
    for k, p in model.named_parameters():
        nz[k] = p.grad is not None
        if p.grad is not None:
            arr["gsum:" + k] = np.array(grad_summary(p.grad, k, seed))
    save(name, shapes, arr, {"B": B, "seed": seed, "has_grad": nz})


def make_ling(R, B, T, seed, name):
    enc = R.encoders.EnhancedTextEncoder({"hidden_dim": 64, "dropout": 0.0})
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 2000, (B, T), generator=g)
    lens = torch.randint(0, T + 1, (B,), generator=g)
    lens[0] = T
    lens[1] = 0
    mask = (torch.arange(T)[None, :] < lens[:, None]).long()
    feats = enc.extract_linguistic_features(ids, mask)
    save(name, {}, {"ids": ids, "mask": mask, "feats": feats}, {"B": B, "T": T, "seed": seed})


def make_ling_wide(R, B, T, seed, name):
    """Realistic vocabulary range (BERT ids up to 30521), repeated tokens, punctuation / special ranges, an all-masked row,
    a mask with holes (not only a prefix)."""
    enc = R.encoders.EnhancedTextEncoder({"hidden_dim": 64, "dropout": 0.0})
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, 30522, (B, T), generator=g)
    small = torch.randint(95, 1040, (B, T), generator=g)
    pick = torch.rand((B, T), generator=g) < 0.5
    ids = torch.where(pick, small, ids)
    rep = torch.rand((B, T), generator=g) < 0.3
    ids = torch.where(rep, ids[:, :1].expand(B, T), ids)
    lens = torch.randint(0, T + 1, (B,), generator=g)
    lens[0] = T
    lens[1] = 0
    lens[2] = 1
    mask = (torch.arange(T)[None, :] < lens[:, None]).long()
    holes = torch.rand((B, T), generator=g) < 0.2
    mask[B // 2:] = mask[B // 2:] * (~holes[B // 2:]).long()
    feats = enc.extract_linguistic_features(ids, mask)
    save(name, {}, {"ids": ids, "mask": mask, "feats": feats}, {"B": B, "T": T, "seed": seed})


def make_metrics(N, D, seed, name, nan_frac=0.0):
    """src/utils/metrics.py run unmodified on float32 arrays (what evaluation.py hands it after .cpu().numpy())."""
    sys.path[:0] = [f"{REF}/src/utils"]
    import metrics as M
    rng = np.random.default_rng(seed)
    tgt = np.tanh(rng.standard_normal((N, D))).astype(np.float32)
    pred = (0.7 * tgt + 0.3 * rng.standard_normal((N, D)) + 0.05).astype(np.float32)
    unc = np.abs(0.4 * rng.standard_normal((N, D)) + 0.3).astype(np.float32)
    if nan_frac > 0:
        pred[rng.random((N, D)) < nan_frac] = np.nan
        unc[rng.random((N, D)) < nan_frac] = np.inf
        unc[rng.random((N, D)) < nan_frac] = np.nan
        # ties in the uncertainties, so quantile edges land exactly on sample values
        unc[: N // 4] = np.round(unc[: N // 4], 1)
    dm = M.DEERMetrics()
    arr = {"pred": pred, "tgt": tgt, "unc": unc}
    arr["ccc"] = np.array([dm.concordance_correlation_coefficient(tgt[:, i], pred[:, i]) for i in range(D)])
    arr["mae"] = np.array([dm.mean_absolute_error(tgt[:, i], pred[:, i]) for i in range(D)])
    arr["rmse"] = np.array([dm.root_mean_squared_error(tgt[:, i], pred[:, i]) for i in range(D)])
    with np.errstate(all="ignore"):
        arr["uce"] = np.array(M.uncertainty_calibration_error(pred, tgt, unc))
        arr["uce5"] = np.array(M.uncertainty_calibration_error(pred, tgt, unc, n_bins=5))
    if nan_frac == 0 and D == 3:
        ev = dm.evaluate_predictions(pred, tgt, unc)
        arr["ev_ccc"] = np.array([ev.ccc_valence, ev.ccc_arousal, ev.ccc_dominance])
        arr["ev_mae"] = np.array([ev.mae_valence, ev.mae_arousal, ev.mae_dominance])
        arr["ev_ece"] = np.array(ev.ece)
        arr["ev_cohens_d"] = np.array([ev.statistical_significance[f"cohens_d_{d}"]
                                       for d in ("valence", "arousal", "dominance")])
    save(name, {}, arr, {"N": N, "D": D, "seed": seed})


def main():
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    if only & {"metrics", "ling_wide"} or "new" in only:
        R = import_reference()
        make_ling_wide(R, 32, 64, 21, "ling_b32")
        make_metrics(1000, 3, 31, "metrics_n1000")
        make_metrics(37, 3, 32, "metrics_n37")
        make_metrics(4099, 3, 33, "metrics_nan", nan_frac=0.02)
        make_metrics(513, 1, 34, "metrics_d1")
        make_metrics(7, 3, 35, "metrics_tiny")
        return
    R = import_reference()
    make_audio(R, 64, 3, 7, 1, "audio_small", True)
    make_audio(R, 512, 2, 12, 2, "audio_full_t12", False)
    make_video(R, 64, 24, 3, 5, 3, "video_small_train", True)
    make_video(R, 64, 24, 3, 5, 3, "video_small_eval", False)
    make_video(R, 64, 24, 3, 1, 4, "video_small_f1", True)
    make_text(R, 64, 3, 6, 5, "text_small")
    make_fusion(R, 64, 48, 40, 64, 32, 4, 5, 6, "fusion_small")
    make_head(R, 64, 32, 6, 7, "head_small")
    make_loss(R, 64, 8, "loss_b64")
    make_loss(R, 1000, 9, "loss_b1000")
    make_loss(R, 3, 10, "loss_b3")
    make_seq_full(R, 4, 300, 50, 64, 11, "seq_full_b4")
    make_pooled(R, 16, 12, "pooled_b16")
    make_ling(R, 6, 16, 13, "ling")


if __name__ == "__main__":
    main()
