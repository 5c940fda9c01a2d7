"""B200 drop-in for /root/reference/src/models/deer.py: DEERLayer (:30-108), DEERLoss (:111-195, the Amini-style
variant) and MultiDimensionalDEER (:198-266).  Output dictionaries carry the reference keys; MultiDimensionalDEER
additionally exposes `gamma/nu/alpha/beta` [B,3] and the raw evidence so the loss can run fused from it."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import ops

NIG_KEYS = ("mu", "nu", "alpha", "beta", "aleatoric_uncertainty", "epistemic_uncertainty", "uncertainty")
EVIDENCE_KEY = "_deer_evidence"  # raw [B,D,4] head output; lets MultiTaskDEERLoss run head+loss fused


def _evidence_net(input_dim, hidden_dim, dropout, out):
    return nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                         nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(), nn.Dropout(dropout),
                         nn.Linear(hidden_dim // 2, out))


class DEERLayer(nn.Module):
    def __init__(self, input_dim: int, output_dim: int = 1, hidden_dim: int = 256, dropout: float = 0.3):
        super().__init__()
        self.input_dim, self.output_dim, self.dropout = input_dim, output_dim, dropout
        self.evidence_net = _evidence_net(input_dim, hidden_dim, dropout, 4 * output_dim)
        for m in self.evidence_net:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0.0)

    def evidence(self, x):
        n = self.evidence_net
        h = ops.linear(x, n[0].weight, n[0].bias, "relu", dropout=self.dropout, training=self.training)
        h = ops.linear(h, n[3].weight, n[3].bias, "relu", dropout=self.dropout, training=self.training)
        return ops.linear(h, n[6].weight, n[6].bias).view(x.shape[0], self.output_dim, 4)

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        nig = ops.nig_head(self.evidence(x))                      # [7,B,output_dim]
        return {k: nig[i] for i, k in enumerate(NIG_KEYS)}


class DEERLoss(nn.Module):
    """deer.py:111-195 (NLL with lgamma(alpha)-lgamma(alpha+1/2), evidence regulariser, clamped KL)."""

    def __init__(self, evidence_weight: float = 1.0, kl_weight: float = 1.0):
        super().__init__()
        self.evidence_weight, self.kl_weight = evidence_weight, kl_weight

    def forward(self, predictions: Dict[str, torch.Tensor], targets: torch.Tensor) -> Dict[str, torch.Tensor]:
        mu, nu, alpha, beta = (predictions[k] for k in ("mu", "nu", "alpha", "beta"))
        if targets.dim() == 1:
            targets = targets.unsqueeze(-1)
        l = ops.amini_loss(mu, nu, alpha, beta, targets.expand_as(mu), self.evidence_weight, self.kl_weight)
        return {"total_loss": l[0], "nll_loss": l[1], "evidence_reg": l[2], "kl_reg": l[3], "mse": l[4]}


class MultiDimensionalDEER(nn.Module):
    def __init__(self, input_dim: int, emotion_dims: int = 3, hidden_dim: int = 256, dropout: float = 0.3):
        super().__init__()
        self.emotion_dims, self.dropout = emotion_dims, dropout
        self.feature_processor = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                                               nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))
        self.deer_heads = nn.ModuleList([DEERLayer(hidden_dim, 1, hidden_dim // 2, dropout)
                                         for _ in range(emotion_dims)])
        self.dimension_names = ["valence", "arousal", "dominance"][:emotion_dims]

    def evidence(self, x: torch.Tensor) -> torch.Tensor:
        """[B,input_dim] -> raw evidence [B,D,4]; the D heads run as grouped GEMMs into one buffer."""
        fp = self.feature_processor
        dr = dict(dropout=self.dropout, training=self.training)    # fused into the GEMM epilogues
        f = ops.linear(x, fp[0].weight, fp[0].bias, "relu", **dr)
        f = ops.linear(f, fp[3].weight, fp[3].bias, "relu", **dr)
        D = self.emotion_dims
        nets = [h.evidence_net for h in self.deer_heads]
        h1 = ops.grouped_linear([f] * D, [n[0].weight for n in nets], [n[0].bias for n in nets], "relu", **dr)
        # (a [B,D,K] tensor as input: head g reads h[:, g]; its gradient comes back as one tensor)
        h2 = ops.grouped_linear(h1, [n[3].weight for n in nets], [n[3].bias for n in nets], "relu", **dr)
        return ops.grouped_linear(h2, [n[6].weight for n in nets], [n[6].bias for n in nets])

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        e = self.evidence(x)
        return nig_dict(e, ops.nig_head(e), self.dimension_names)


def nig_dict(evidence, nig, names, trailing_dim: bool = True) -> Dict[str, torch.Tensor]:
    """Assemble the reference's prediction dictionary from the stacked NIG outputs nig [7,B,D]."""
    out = {}
    for i, d in enumerate(names):
        for j, k in enumerate(NIG_KEYS):
            out[f"{d}_{k}"] = nig[j][:, i:i + 1] if trailing_dim else nig[j][:, i]
    out["mu_all"] = nig[0]
    out["uncertainty_all"] = nig[6]
    out["gamma"], out["nu"], out["alpha"], out["beta"] = nig[0], nig[1], nig[2], nig[3]
    out[EVIDENCE_KEY] = evidence
    return out


def test_deer_implementation():
    """Smoke check the reference driver imports (run_multimodal_deer.py:76); needs a CUDA device."""
    head = MultiDimensionalDEER(512).cuda()
    out = head(torch.randn(16, 512, device="cuda"))
    assert out["mu_all"].shape == (16, 3)
    return True
