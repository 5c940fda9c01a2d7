"""CPU baseline port of the reference's sequence path, built from the SAME stock torch.nn layers the reference calls
(nn.LSTM -> oneDNN, nn.MultiheadAttention, nn.Conv1d, nn.BatchNorm1d, nn.LayerNorm, F.softplus, torch.lgamma), wired
as the reference wires them.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py `cpu_baseline` and `--impl reference`):
/root/reference is a Python repo that does not travel to the GPU box, so this port is what is timed on the box's
host cores ("kind": "port").  `tests/test_oracle_golden.py::test_torch_baseline_matches_oracle` checks it against the
golden-pinned oracle, so the timed code computes the reference's numbers.

Reference lines: encoders.py:82-107,356-389 (audio), :443-475,531-548 (video), :597-625,733-761 (text);
fusion.py:35-343; deer.py:30-108,198-266; losses.py:40-348.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

DIMS = ("valence", "arousal", "dominance")


def _scorer(d):
    return nn.Sequential(nn.Linear(d, d // 2), nn.Tanh(), nn.Linear(d // 2, 1), nn.Softmax(dim=1))


class AudioEnc(nn.Module):
    def __init__(self, d=512, p=0.3):
        super().__init__()
        self.lstm = nn.LSTM(84, d // 2, num_layers=2, batch_first=True, dropout=p, bidirectional=True)
        self.attention = _scorer(d)
        self.output_projection = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Dropout(p), nn.Linear(d, d), nn.LayerNorm(d))

    def forward(self, x):
        h, _ = self.lstm(x)
        w = self.attention(h)
        return self.output_projection(torch.sum(h * w, dim=1))


class VideoEnc(nn.Module):
    def __init__(self, din=256, d=512, p=0.3):
        super().__init__()
        self.spatial_projection = nn.Sequential(nn.Linear(din, d), nn.ReLU(), nn.Dropout(p))
        self.temporal_cnn = nn.Sequential(nn.Conv1d(d, d, 3, padding=1), nn.BatchNorm1d(d), nn.ReLU(), nn.Dropout(p),
                                          nn.Conv1d(d, d, 3, padding=1), nn.BatchNorm1d(d), nn.ReLU())
        self.temporal_attention = _scorer(d)
        self.output_projection = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Dropout(p), nn.LayerNorm(d))

    def forward(self, x):
        h = self.spatial_projection(x)
        h = self.temporal_cnn(h.transpose(1, 2)).transpose(1, 2)
        w = self.temporal_attention(h)
        return self.output_projection(torch.sum(h * w, dim=1))


class TextEnc(nn.Module):
    def __init__(self, d=512, p=0.3):
        super().__init__()
        self.token_attention = _scorer(768)
        self.bert_projection = nn.Sequential(nn.Linear(768, d), nn.ReLU(), nn.Dropout(p))
        self.linguistic_projection = nn.Sequential(nn.Linear(10, d // 4), nn.ReLU(), nn.Dropout(p))
        self.output_projection = nn.Sequential(nn.Linear(d + d // 4, d), nn.ReLU(), nn.Dropout(p), nn.LayerNorm(d))

    def forward(self, tok, mask, ling):
        m = mask.unsqueeze(-1)
        x = tok * m
        w = self.token_attention(x) * m
        w = w / (torch.sum(w, dim=1, keepdim=True) + 1e-10)
        agg = torch.sum(x * w, dim=1)
        return self.output_projection(torch.cat([self.bert_projection(agg), self.linguistic_projection(ling)], dim=1))


class AVFusion(nn.Module):
    def __init__(self, d, e, heads, p):
        super().__init__()
        self.audio_projection, self.video_projection = nn.Linear(d, e), nn.Linear(d, e)
        self.cross_attention = nn.MultiheadAttention(e, heads, dropout=p, batch_first=True)
        self.fusion_layers = nn.Sequential(nn.Linear(2 * e, e), nn.ReLU(), nn.Dropout(p), nn.LayerNorm(e))

    def forward(self, a, v):
        a, v = self.audio_projection(a).unsqueeze(1), self.video_projection(v).unsqueeze(1)
        aa, _ = self.cross_attention(query=a, key=v, value=v)
        va, _ = self.cross_attention(query=v, key=a, value=a)
        return self.fusion_layers(torch.cat([aa.squeeze(1), va.squeeze(1)], dim=-1))


class TriFusion(nn.Module):
    def __init__(self, e, d, heads, p):
        super().__init__()
        self.audiovisual_projection, self.text_projection = nn.Linear(e, d), nn.Linear(d, d)
        self.modality_attention = nn.MultiheadAttention(d, heads, dropout=p, batch_first=True)
        self.final_fusion = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Dropout(p), nn.LayerNorm(d))

    def forward(self, av, t):
        m = torch.stack([self.audiovisual_projection(av), self.text_projection(t)], dim=1)
        o, _ = self.modality_attention(query=m, key=m, value=m)
        return self.final_fusion(o.mean(dim=1))


class Fusion(nn.Module):
    def __init__(self, d=512, heads=8, p=0.3):
        super().__init__()
        self.audio_visual_fusion = AVFusion(d, d // 2, heads, p)
        self.trimodal_fusion = TriFusion(d // 2, d, heads, p)
        self.output_projection = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Dropout(p), nn.LayerNorm(d))

    def forward(self, a, v, t):
        return self.output_projection(self.trimodal_fusion(self.audio_visual_fusion(a, v), t))


class Head(nn.Module):
    def __init__(self, d=512, h=256, p=0.3):
        super().__init__()
        self.feature_processor = nn.Sequential(nn.Linear(d, h), nn.ReLU(), nn.Dropout(p), nn.Linear(h, h), nn.ReLU(),
                                               nn.Dropout(p))
        self.deer_heads = nn.ModuleList([nn.ModuleDict({"evidence_net": nn.Sequential(
            nn.Linear(h, h // 2), nn.ReLU(), nn.Dropout(p), nn.Linear(h // 2, h // 4), nn.ReLU(), nn.Dropout(p),
            nn.Linear(h // 4, 4))}) for _ in range(3)])

    def forward(self, x):
        f = self.feature_processor(x)
        out = {}
        for i, d in enumerate(DIMS):
            e = self.deer_heads[i]["evidence_net"](f).view(x.shape[0], 1, 4)
            out[f"{d}_mu"] = e[:, :, 0]
            out[f"{d}_nu"] = F.softplus(e[:, :, 1]) + 1e-6
            out[f"{d}_alpha"] = F.softplus(e[:, :, 2]) + 1.0
            out[f"{d}_beta"] = F.softplus(e[:, :, 3]) + 1e-6
        return out


def deer_loss(g, nu, al, be, y, eps=1e-8):
    err = y - g
    lp = (0.5 * torch.log(nu / (2 * math.pi + eps)) + al * torch.log(be + eps) - torch.lgamma(al + eps)
          - (al + 0.5) * torch.log(be + 0.5 * nu * err.pow(2) + eps))
    nll = -torch.mean(lp)
    a = torch.abs(err)
    reg = torch.mean(a.pow(2) * (2 * be + nu * a.pow(2)))
    kl = torch.mean((al - 1).pow(2)) + 0.1 * torch.mean((torch.log(be + eps) - math.log(1 + eps)) ** 2)
    conf = (1.0 / (1.0 + be / (al - 1 + eps))).flatten()
    ef = a.flatten()
    edges = torch.linspace(0, 1, 11)
    ece = 0.0
    for k in range(10):
        inb = (conf > edges[k]) & (conf <= edges[k + 1])
        if inb.sum() > 0:                                     # host sync per bin, as in losses.py:217
            ece = ece + inb.sum().float() / conf.size(0) * torch.abs(conf[inb].mean() - (1.0 - ef[inb].mean()))
    return nll + 0.1 * reg + 0.01 * kl + 0.05 * ece


def multitask_loss(pred: Dict[str, torch.Tensor], y):
    total, us = 0.0, []
    for i, d in enumerate(DIMS):
        total = total + deer_loss(pred[f"{d}_mu"], pred[f"{d}_nu"], pred[f"{d}_alpha"], pred[f"{d}_beta"], y[:, i:i + 1])
        us.append((pred[f"{d}_beta"] / (pred[f"{d}_alpha"] - 1 + 1e-8)).mean(dim=0))
    cd = sum(F.mse_loss(us[i], us[j]) for i in range(3) for j in range(i + 1, 3)) / 3
    return (total + 0.05 * cd) / 3


class SequenceBaseline(nn.Module):
    def __init__(self, dropout=0.3):
        super().__init__()
        self.audio_encoder, self.video_encoder, self.text_encoder = AudioEnc(p=dropout), VideoEnc(p=dropout), TextEnc(p=dropout)
        self.fusion, self.deer = Fusion(p=dropout), Head(p=dropout)

    def forward(self, audio, video, text, mask, ling):
        return self.deer(self.fusion(self.audio_encoder(audio), self.video_encoder(video),
                                     self.text_encoder(text, mask, ling)))


def time_cpu_baseline(mode: str, batch: int, iters: int = 3, warmup: int = 1, dropout: float = 0.3, seed: int = 42):
    """Time the CPU port: returns (samples_per_s, seconds_per_step, threads).  mode: 'train' or 'infer'."""
    import os
    import time
    torch.manual_seed(seed)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = SequenceBaseline(dropout)
    B = batch
    audio, video, text = torch.randn(B, 300, 84), torch.randn(B, 50, 256), torch.randn(B, 64, 768)
    mask, ling = torch.ones(B, 64), torch.zeros(B, 10)
    y = torch.tanh(torch.randn(B, 3) + 0.1 * torch.randn(B, 3))
    if mode == "train":
        model.train()
        params = [p for p in model.parameters()]
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-5, eps=1e-8)

        def step():
            opt.zero_grad()
            loss = multitask_loss(model(audio, video, text, mask, ling), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return float(loss)
    else:
        model.eval()

        def step():
            with torch.no_grad():
                return float(model(audio, video, text, mask, ling)["valence_mu"].sum())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = (time.perf_counter() - t0) / iters
    return B / dt, dt, torch.get_num_threads()
