"""CPU oracle for the DEER forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch-CPU tensor algebra, fp64 or
fp32, no nn.Module, no nn.LSTM / nn.MultiheadAttention / F.layer_norm calls) of
the reference algorithm for the path named by BASELINE.json `north_star`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product package never does.

Parity status: the reference repo holds no golden vectors or known-answer tests
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, executed in the build container by
`tests/golden/make_golden.py` and committed under `tests/golden/*.npz`
(`tests/test_oracle_golden.py` checks every function here against them).

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  Weights are passed as a flat dict with the reference's
own state_dict key names, so one state_dict drives reference, oracle and CUDA.

Gradients: obtained with torch autograd over these primitive tensor ops (the
reference itself uses plain autograd).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

DIMS = ("valence", "arousal", "dominance")


# --------------------------------------------------------------------------- #
# primitives (torch semantics copied per SURVEY.md section 8c)
# --------------------------------------------------------------------------- #
def linear(x, w, b=None):
    """nn.Linear: y = x W^T + b."""
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x, g, b, eps=1e-5):
    """nn.LayerNorm over the last dim: biased variance, eps inside the sqrt."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def softplus(x):
    """F.softplus with beta=1, threshold=20 (linear branch above 20)."""
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def softmax(x, dim):
    m = x.max(dim=dim, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=dim, keepdim=True)


# --------------------------------------------------------------------------- #
# A1: 2-layer bidirectional LSTM  (encoders.py:82-89, call :380)
# --------------------------------------------------------------------------- #
def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.LSTM layer, zero initial state.
    x [B,T,I]; gate order i,f,g,o; c_t=f*c+i*g; h_t=o*tanh(c_t)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    # (unbind, not pre[:, t]: the backward of T selects would materialise T zero-filled [B,T,4H] tensors)
    pre = linear(x, w_ih, b_ih + b_hh).unbind(1)  # T x [B,4H]
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    w_hh_t = w_hh.t()
    for t in order:
        g = pre[t] + h @ w_hh_t
        i_, f_, g_, o_ = g.split(H, dim=1)
        i_, f_, o_ = torch.sigmoid(i_), torch.sigmoid(f_), torch.sigmoid(o_)
        g_ = torch.tanh(g_)
        c = f_ * c + i_ * g_
        h = o_ * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm(x, sd, prefix="lstm.", num_layers=2):
    """nn.LSTM(bidirectional, batch_first) with inter-layer dropout disabled
    (parity runs use dropout=0, SURVEY.md section 7 hard part 5)."""
    out = x
    for l in range(num_layers):
        f = lstm_direction(out, sd[f"{prefix}weight_ih_l{l}"], sd[f"{prefix}weight_hh_l{l}"],
                           sd[f"{prefix}bias_ih_l{l}"], sd[f"{prefix}bias_hh_l{l}"], False)
        r = lstm_direction(out, sd[f"{prefix}weight_ih_l{l}_reverse"], sd[f"{prefix}weight_hh_l{l}_reverse"],
                           sd[f"{prefix}bias_ih_l{l}_reverse"], sd[f"{prefix}bias_hh_l{l}_reverse"], True)
        out = torch.cat([f, r], dim=-1)
    return out


def attn_pool(x, w1, b1, w2, b2):
    """Linear-Tanh-Linear-Softmax(dim=1) scorer and weighted sum over time
    (encoders.py:93-98,383-384; same block at :462-467,:543-544)."""
    s = linear(torch.tanh(linear(x, w1, b1)), w2, b2)  # [B,T,1]
    w = softmax(s, dim=1)
    return (x * w).sum(dim=1), w


# --------------------------------------------------------------------------- #
# A1-A3: EnhancedAudioEncoder.forward on pre-extracted features (encoders.py:356-389)
# --------------------------------------------------------------------------- #
def audio_encoder(x, sd, p=""):
    h = bilstm(x, sd, p + "lstm.")
    pooled, _ = attn_pool(h, sd[p + "attention.0.weight"], sd[p + "attention.0.bias"],
                          sd[p + "attention.2.weight"], sd[p + "attention.2.bias"])
    y = torch.relu(linear(pooled, sd[p + "output_projection.0.weight"], sd[p + "output_projection.0.bias"]))
    y = linear(y, sd[p + "output_projection.3.weight"], sd[p + "output_projection.3.bias"])
    return layer_norm(y, sd[p + "output_projection.4.weight"], sd[p + "output_projection.4.bias"])


# --------------------------------------------------------------------------- #
# V: EnhancedVideoEncoder post-backbone half (encoders.py:443-475, forward :531-548)
# --------------------------------------------------------------------------- #
def conv1d_k3(x, w, b):
    """nn.Conv1d(C,C,kernel_size=3,padding=1) on channels-last x [B,T,Cin];
    w [Cout,Cin,3]."""
    B, T, C = x.shape
    z = x.new_zeros(B, 1, C)
    xp = torch.cat([z, x, z], dim=1)
    y = b
    for k in range(3):
        y = y + xp[:, k:k + T] @ w[:, :, k].t()
    return y


def batchnorm_cl(x, g, b, running_mean, running_var, training, eps=1e-5):
    """nn.BatchNorm1d over channels of channels-last x [B,T,C]: batch statistics
    (biased variance) in training, running statistics in eval."""
    if training:
        mu = x.mean(dim=(0, 1))
        var = ((x - mu) ** 2).mean(dim=(0, 1))
    else:
        mu, var = running_mean, running_var
    return (x - mu) / torch.sqrt(var + eps) * g + b


def video_encoder(x, sd, p="", training=True):
    """x [B,F,Din] frame features (spatial backbone bypassed)."""
    h = torch.relu(linear(x, sd[p + "spatial_projection.0.weight"], sd[p + "spatial_projection.0.bias"]))
    if x.shape[1] > 1:
        h = conv1d_k3(h, sd[p + "temporal_cnn.0.weight"], sd[p + "temporal_cnn.0.bias"])
        h = torch.relu(batchnorm_cl(h, sd[p + "temporal_cnn.1.weight"], sd[p + "temporal_cnn.1.bias"],
                                    sd[p + "temporal_cnn.1.running_mean"], sd[p + "temporal_cnn.1.running_var"], training))
        h = conv1d_k3(h, sd[p + "temporal_cnn.4.weight"], sd[p + "temporal_cnn.4.bias"])
        h = torch.relu(batchnorm_cl(h, sd[p + "temporal_cnn.5.weight"], sd[p + "temporal_cnn.5.bias"],
                                    sd[p + "temporal_cnn.5.running_mean"], sd[p + "temporal_cnn.5.running_var"], training))
        pooled, _ = attn_pool(h, sd[p + "temporal_attention.0.weight"], sd[p + "temporal_attention.0.bias"],
                              sd[p + "temporal_attention.2.weight"], sd[p + "temporal_attention.2.bias"])
    else:
        pooled = h[:, 0]
    y = torch.relu(linear(pooled, sd[p + "output_projection.0.weight"], sd[p + "output_projection.0.bias"]))
    return layer_norm(y, sd[p + "output_projection.3.weight"], sd[p + "output_projection.3.bias"])


# --------------------------------------------------------------------------- #
# T1: EnhancedTextEncoder post-BERT half (encoders.py:597-625, forward :733-761)
# --------------------------------------------------------------------------- #
def text_encoder(tok, mask, ling, sd, p=""):
    """tok [B,T,768] token embeddings, mask [B,T] (float 0/1), ling [B,10]."""
    m = mask.unsqueeze(-1)
    x = tok * m                                                        # :734-735
    s = linear(torch.tanh(linear(x, sd[p + "token_attention.0.weight"], sd[p + "token_attention.0.bias"])),
               sd[p + "token_attention.2.weight"], sd[p + "token_attention.2.bias"])
    w = softmax(s, dim=1) * m                                          # :738-739 (softmax over ALL T, then mask)
    w = w / (w.sum(dim=1, keepdim=True) + 1e-10)                       # :742-743
    agg = (x * w).sum(dim=1)                                           # :746
    pb = torch.relu(linear(agg, sd[p + "bert_projection.0.weight"], sd[p + "bert_projection.0.bias"]))
    pl = torch.relu(linear(ling, sd[p + "linguistic_projection.0.weight"], sd[p + "linguistic_projection.0.bias"]))
    y = torch.relu(linear(torch.cat([pb, pl], dim=1), sd[p + "output_projection.0.weight"],
                          sd[p + "output_projection.0.bias"]))
    return layer_norm(y, sd[p + "output_projection.3.weight"], sd[p + "output_projection.3.bias"])


def linguistic_features(ids, mask):
    """T2: EnhancedTextEncoder.extract_linguistic_features (encoders.py:648-699),
    integer statistics of the valid token ids; returns float32 [B,10]."""
    B = ids.shape[0]
    out = torch.zeros(B, 10, dtype=torch.float32)
    for b in range(B):
        valid = ids[b][mask[b].bool()]
        n = int(valid.numel())
        if n == 0:
            continue
        uniq, counts = torch.unique(valid, return_counts=True)
        out[b, 0] = n / 128.0
        out[b, 1] = uniq.numel() / n
        bc = torch.bincount(valid).float()
        out[b, 2] = bc.mean()
        out[b, 3] = bc.max()
        out[b, 4] = ((valid >= 999) & (valid <= 1030)).float().mean()
        out[b, 5] = ((valid >= 100) & (valid <= 999)).float().mean()
    return out


# --------------------------------------------------------------------------- #
# F1-F3: HierarchicalMultimodalFusion, uncertainties=None (fusion.py:119-171)
# --------------------------------------------------------------------------- #
def mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, heads):
    """nn.MultiheadAttention(batch_first=True), dropout 0, need_weights=True with
    head-averaged weights.  q_in [B,Lq,E], k_in/v_in [B,Lk,E]."""
    B, Lq, E = q_in.shape
    Lk = k_in.shape[1]
    d = E // heads
    q = linear(q_in, in_w[:E], in_b[:E]).view(B, Lq, heads, d).transpose(1, 2)
    k = linear(k_in, in_w[E:2 * E], in_b[E:2 * E]).view(B, Lk, heads, d).transpose(1, 2)
    v = linear(v_in, in_w[2 * E:], in_b[2 * E:]).view(B, Lk, heads, d).transpose(1, 2)
    p = softmax((q / math.sqrt(d)) @ k.transpose(-1, -2), dim=-1)      # [B,h,Lq,Lk]
    o = (p @ v).transpose(1, 2).reshape(B, Lq, E)
    return linear(o, out_w, out_b), p.mean(dim=1)


def audio_visual_fusion(a, v, sd, p, heads):
    """fusion.py:223-271."""
    ap = linear(a, sd[p + "audio_projection.weight"], sd[p + "audio_projection.bias"]).unsqueeze(1)
    vp = linear(v, sd[p + "video_projection.weight"], sd[p + "video_projection.bias"]).unsqueeze(1)
    mw = (sd[p + "cross_attention.in_proj_weight"], sd[p + "cross_attention.in_proj_bias"],
          sd[p + "cross_attention.out_proj.weight"], sd[p + "cross_attention.out_proj.bias"])
    aa, wa = mha(ap, vp, vp, *mw, heads)
    va, wv = mha(vp, ap, ap, *mw, heads)
    cat = torch.cat([aa.squeeze(1), va.squeeze(1)], dim=-1)
    y = torch.relu(linear(cat, sd[p + "fusion_layers.0.weight"], sd[p + "fusion_layers.0.bias"]))
    y = layer_norm(y, sd[p + "fusion_layers.3.weight"], sd[p + "fusion_layers.3.bias"])
    return y, {"audio_to_video": wa.squeeze(1), "video_to_audio": wv.squeeze(1)}


def trimodal_fusion(av, t, sd, p, heads):
    """fusion.py:308-343."""
    avp = linear(av, sd[p + "audiovisual_projection.weight"], sd[p + "audiovisual_projection.bias"])
    tp = linear(t, sd[p + "text_projection.weight"], sd[p + "text_projection.bias"])
    m = torch.stack([avp, tp], dim=1)
    o, w = mha(m, m, m, sd[p + "modality_attention.in_proj_weight"], sd[p + "modality_attention.in_proj_bias"],
               sd[p + "modality_attention.out_proj.weight"], sd[p + "modality_attention.out_proj.bias"], heads)
    pooled = o.mean(dim=1)
    y = torch.relu(linear(pooled, sd[p + "final_fusion.0.weight"], sd[p + "final_fusion.0.bias"]))
    return layer_norm(y, sd[p + "final_fusion.3.weight"], sd[p + "final_fusion.3.bias"]), w


def hierarchical_fusion(a, v, t, sd, p="", heads=8):
    av, avw = audio_visual_fusion(a, v, sd, p + "audio_visual_fusion.", heads)
    tri, triw = trimodal_fusion(av, t, sd, p + "trimodal_fusion.", heads)
    y = torch.relu(linear(tri, sd[p + "output_projection.0.weight"], sd[p + "output_projection.0.bias"]))
    y = layer_norm(y, sd[p + "output_projection.3.weight"], sd[p + "output_projection.3.bias"])
    return {"fused_features": y, "audiovisual_features": av, "trimodal_features": tri,
            "av_attention_weights": avw, "trimodal_attention_weights": triw, "uncertainty_weights": None}


# --------------------------------------------------------------------------- #
# H1/H2: NIG heads (deer.py:30-108,198-266; complete_project.py:369-417)
# --------------------------------------------------------------------------- #
def nig_from_evidence(e):
    """e [...,4] raw evidence -> dict (deer.py:90-98)."""
    mu = e[..., 0]
    nu = softplus(e[..., 1]) + 1e-6
    alpha = softplus(e[..., 2]) + 1.0
    beta = softplus(e[..., 3]) + 1e-6
    alea = beta / (alpha - 1)
    epis = beta / (nu * (alpha - 1))
    return {"mu": mu, "nu": nu, "alpha": alpha, "beta": beta,
            "aleatoric_uncertainty": alea, "epistemic_uncertainty": epis, "uncertainty": alea + epis}


def multidim_deer(x, sd, p=""):
    """MultiDimensionalDEER.forward (deer.py:233-266), dropout 0."""
    f = torch.relu(linear(x, sd[p + "feature_processor.0.weight"], sd[p + "feature_processor.0.bias"]))
    f = torch.relu(linear(f, sd[p + "feature_processor.3.weight"], sd[p + "feature_processor.3.bias"]))
    out = {}
    for i, d in enumerate(DIMS):
        q = f"{p}deer_heads.{i}.evidence_net."
        h = torch.relu(linear(f, sd[q + "0.weight"], sd[q + "0.bias"]))
        h = torch.relu(linear(h, sd[q + "3.weight"], sd[q + "3.bias"]))
        e = linear(h, sd[q + "6.weight"], sd[q + "6.bias"]).view(x.shape[0], 1, 4)
        for k, v in nig_from_evidence(e).items():
            out[f"{d}_{k}"] = v
    out["mu_all"] = torch.cat([out[f"{d}_mu"] for d in DIMS], dim=1)
    out["uncertainty_all"] = torch.cat([out[f"{d}_uncertainty"] for d in DIMS], dim=1)
    return out


# --------------------------------------------------------------------------- #
# L1/L2: losses.DEERLoss / MultiTaskDEERLoss (losses.py:40-348)
# --------------------------------------------------------------------------- #
def ece_bin_edges(n_bins=10, dtype=torch.float32):
    """torch.linspace(0,1,n_bins+1) as the reference builds it (losses.py:207)."""
    return torch.linspace(0, 1, n_bins + 1, dtype=dtype)


def deer_loss(gamma, nu, alpha, beta, y, reg_weight=0.1, kl_weight=0.01, ece_weight=0.05, eps=1e-8):
    """losses.DEERLoss.forward (losses.py:72-226) on [B,1] tensors."""
    err = y - gamma
    lp = (0.5 * torch.log(nu / (2 * math.pi + eps)) + alpha * torch.log(beta + eps)
          - torch.lgamma(alpha + eps) - (alpha + 0.5) * torch.log(beta + 0.5 * nu * err ** 2 + eps))
    nll = -lp.mean()                                                     # :141-151
    a = err.abs()
    reg = (a ** 2 * (2 * beta + nu * a ** 2)).mean()                     # :165-167
    kl = ((alpha - 1.0) ** 2).mean() + 0.1 * ((torch.log(beta + eps) - math.log(1.0 + eps)) ** 2).mean()  # :178-185
    ece = gamma.new_zeros(())
    if ece_weight > 0:                                                   # :196-226
        u = beta / (alpha - 1 + eps)
        conf = (1.0 / (1.0 + u)).flatten()
        ef = a.flatten()
        edges = ece_bin_edges(10, torch.float32).to(conf.dtype)
        n = conf.numel()
        for k in range(10):
            inb = (conf > edges[k]) & (conf <= edges[k + 1])
            cnt = int(inb.sum())
            if cnt > 0:
                # `in_bin.sum().float() / N` (losses.py:222): the bin weight is rounded to float32
                # even when the module runs in float64.
                wgt = (torch.tensor(cnt, dtype=torch.float32) / n).to(conf.dtype)
                ece = ece + wgt * (conf[inb].mean() - (1.0 - ef[inb].mean())).abs()
    total = nll + reg_weight * reg + kl_weight * kl + ece_weight * ece
    return {"total_loss": total, "nll_loss": nll, "reg_loss": reg, "kl_loss": kl, "ece_loss": ece,
            "batch_size": gamma.shape[0]}


def multitask_deer_loss(pred: Dict[str, torch.Tensor], y, cross_dim_weight=0.05, task_weights=None, **kw):
    """losses.MultiTaskDEERLoss.forward (losses.py:268-348)."""
    out = {}
    total = 0.0
    us = []
    for i, d in enumerate(DIMS):
        g = pred.get(f"{d}_gamma", pred.get(f"{d}_mu"))
        nu = pred.get(f"{d}_nu", pred.get(f"{d}_lambda"))
        al, be = pred[f"{d}_alpha"], pred[f"{d}_beta"]
        r = deer_loss(g, nu, al, be, y[:, i:i + 1], **kw)
        w = 1.0 if task_weights is None else task_weights[d]
        total = total + w * r["total_loss"]
        for k, v in r.items():
            out[f"{d}_{k}"] = v
        us.append((be / (al - 1 + 1e-8)).mean(dim=0))
    if cross_dim_weight > 0:
        cd = 0.0
        for i in range(3):
            for j in range(i + 1, 3):
                cd = cd + ((us[i] - us[j]) ** 2).mean()
        cd = cd / 3
        total = total + cross_dim_weight * cd
        out["cross_dim_loss"] = cd
    out["total_loss"] = total / 3
    return out


def amini_deer_loss(mu, nu, alpha, beta, y, evidence_weight=1.0, kl_weight=1.0):
    """L3: deer.DEERLoss.forward (deer.py:125-195)."""
    if y.dim() == 1:
        y = y.unsqueeze(-1)
    se = (y - mu) ** 2
    nll = (0.5 * torch.log(math.pi / nu) - alpha * torch.log(2 * beta) + torch.lgamma(alpha)
           - torch.lgamma(alpha + 0.5) + (alpha + 0.5) * torch.log(beta + nu * se / 2))
    reg = (nu * se + 2 * beta * (1 + nu)) / (2 * nu * (1 + nu))
    kl = torch.clamp(0.5 * (nu - 1) + alpha * torch.log(beta) - torch.lgamma(alpha) + torch.lgamma(alpha + 0.5)
                     - 0.5 * torch.log(2 * math.pi * beta), min=0)
    return {"total_loss": nll.mean() + evidence_weight * reg.mean() + kl_weight * kl.mean(),
            "nll_loss": nll.mean(), "evidence_reg": reg.mean(), "kl_reg": kl.mean(), "mse": se.mean()}


# --------------------------------------------------------------------------- #
# Sequence composite and pooled model
# --------------------------------------------------------------------------- #
def sequence_model(audio, video, text, mask, ling, sd, training=True):
    """audio [B,Ta,84], video [B,Tv,256], text [B,Tt,768], mask [B,Tt], ling [B,10]."""
    a = audio_encoder(audio, sd, "audio_encoder.")
    v = video_encoder(video, sd, "video_encoder.", training)
    t = text_encoder(text, mask, ling, sd, "text_encoder.")
    fus = hierarchical_fusion(a, v, t, sd, "fusion.")
    out = multidim_deer(fus["fused_features"], sd, "deer.")
    out["fused_features"] = fus["fused_features"]
    out["audio_encoded"], out["video_encoded"], out["text_encoded"] = a, v, t
    return out


def sequence_model_loss(audio, video, text, mask, ling, y, sd, training=True):
    out = sequence_model(audio, video, text, mask, ling, sd, training)
    return out, multitask_deer_loss(out, y)


# ---- P: CompleteDEERModel (complete_project.py:61-602), dropout 0 ---------- #
def _pooled_encoder(x, sd, p, layers=3):
    """EnhancedModalityEncoder.forward (complete_project.py:99-118)."""
    h = torch.relu(linear(x, sd[p + "input_projection.0.weight"], sd[p + "input_projection.0.bias"]))
    h = layer_norm(h, sd[p + "input_projection.2.weight"], sd[p + "input_projection.2.bias"])
    for l in range(layers):
        q = f"{p}encoder_layers.{l}.layers."
        r = torch.relu(linear(h, sd[q + "0.weight"], sd[q + "0.bias"]))
        h = h + layer_norm(r, sd[q + "3.weight"], sd[q + "3.bias"])
    return linear(h, sd[p + "output_projection.weight"], sd[p + "output_projection.bias"])


def _seq1_attention(x_kv, sd, p):
    """MultiHeadAttention.forward with one key (complete_project.py:145-183):
    softmax over a single key is identically 1, so out = Wo(Wv x + bv) + bo."""
    v = linear(x_kv, sd[p + "value_proj.weight"], sd[p + "value_proj.bias"])
    return linear(v, sd[p + "output_proj.weight"], sd[p + "output_proj.bias"])


def _uncert(x, sd, p):
    """UncertaintyEstimator (complete_project.py:186-212)."""
    h = torch.relu(linear(x, sd[p + "estimator.0.weight"], sd[p + "estimator.0.bias"]))
    h = torch.relu(linear(h, sd[p + "estimator.3.weight"], sd[p + "estimator.3.bias"]))
    return torch.sigmoid(linear(h, sd[p + "estimator.5.weight"], sd[p + "estimator.5.bias"]))


def pooled_model(audio, video, text, sd, layers=3):
    """CompleteDEERModel.forward (complete_project.py:518-588)."""
    a = _pooled_encoder(audio, sd, "audio_encoder.", layers)
    v = _pooled_encoder(video, sd, "video_encoder.", layers)
    t = _pooled_encoder(text, sd, "text_encoder.", layers)
    p = "attention_module."
    ua, uv, ut = (_uncert(z, sd, p + "uncertainty_estimator.") for z in (a, v, t))
    sa, sv, st = (_seq1_attention(z, sd, p + "self_attention.") for z in (a, v, t))
    ca, cv, ct = (_seq1_attention(z, sd, p + "cross_attention.") for z in (a, v, t))
    wi = torch.cat([sa, sv, st, ua, uv, ut], dim=1)
    w = torch.relu(linear(wi, sd[p + "weight_network.0.weight"], sd[p + "weight_network.0.bias"]))
    w = softmax(linear(w, sd[p + "weight_network.3.weight"], sd[p + "weight_network.3.bias"]), dim=1)
    af = w[:, 0:1] * sa + (1 - ua) * ca
    vf = w[:, 1:2] * sv + (1 - uv) * cv
    tf = w[:, 2:3] * st + (1 - ut) * ct
    p = "fusion_module."
    av = torch.relu(linear(torch.cat([af, vf], 1), sd[p + "av_fusion.0.weight"], sd[p + "av_fusion.0.bias"]))
    av = layer_norm(av, sd[p + "av_fusion.3.weight"], sd[p + "av_fusion.3.bias"])
    av = torch.relu(linear(av, sd[p + "av_fusion.4.weight"], sd[p + "av_fusion.4.bias"]))
    tc = torch.cat([av, tf], 1)
    gate = torch.sigmoid(linear(tc, sd[p + "fusion_gate.0.weight"], sd[p + "fusion_gate.0.bias"]))
    tr = torch.relu(linear(tc, sd[p + "trimodal_fusion.0.weight"], sd[p + "trimodal_fusion.0.bias"]))
    tr = layer_norm(tr, sd[p + "trimodal_fusion.3.weight"], sd[p + "trimodal_fusion.3.bias"])
    tr = torch.relu(linear(tr, sd[p + "trimodal_fusion.4.weight"], sd[p + "trimodal_fusion.4.bias"]))
    fused = gate * tr + (1 - gate) * av
    out = {}
    for d in DIMS:
        q = f"prediction_heads.{d}.evidence_network."
        h = torch.relu(linear(fused, sd[q + "0.weight"], sd[q + "0.bias"]))
        h = torch.relu(linear(h, sd[q + "3.weight"], sd[q + "3.bias"]))
        e = linear(h, sd[q + "6.weight"], sd[q + "6.bias"])
        for k, val in nig_from_evidence(e).items():
            out[f"{d}_{k}"] = val
    out["mu_all"] = torch.stack([out[f"{d}_mu"] for d in DIMS], dim=1)
    out["uncertainty_all"] = torch.stack([out[f"{d}_uncertainty"] for d in DIMS], dim=1)
    sc = out["uncertainty_all"] / sd["calibration_layer.temperature"].unsqueeze(0)
    q = "calibration_layer.calibration_network."
    cal = []
    for i in range(3):
        h = torch.relu(linear(sc[:, i:i + 1], sd[q + "0.weight"], sd[q + "0.bias"]))
        h = torch.relu(linear(h, sd[q + "2.weight"], sd[q + "2.bias"]))
        cal.append(torch.sigmoid(linear(h, sd[q + "4.weight"], sd[q + "4.bias"])))
    out["calibrated_uncertainty"] = torch.cat(cal, dim=1)
    out["attention_weights"] = w
    out["modality_uncertainties"] = torch.cat([ua, uv, ut], dim=1)
    out["fused_features"] = fused
    return out


def pooled_loss_inputs(out):
    """The pooled heads emit 1-D [B] tensors; losses.DEERLoss unsqueezes them
    (losses.py:98-104).  Returns a dict shaped like the sequence heads'."""
    r = {}
    for d in DIMS:
        for k in ("mu", "nu", "alpha", "beta"):
            r[f"{d}_{k}"] = out[f"{d}_{k}"].unsqueeze(-1)
    return r
