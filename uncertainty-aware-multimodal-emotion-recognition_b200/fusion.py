"""B200 drop-in for HierarchicalMultimodalFusion (/root/reference/src/models/fusion.py:35-343).

Same constructor signatures, state_dict keys (incl. nn.MultiheadAttention's packed in_proj_weight/in_proj_bias and
the never-executed uncertainty_gate.* parameters) and output dictionary.  nn.MultiheadAttention objects are parameter
containers only.  Algebra used (verified against the reference in tests/golden):
  * sequence length 1 (AudioVisualFusion, fusion.py:244-255): softmax over one key == 1, so
    attended = out_proj(V-rows of in_proj applied to the key/value token); Q,K rows get exactly zero gradient.
  * TrimodalFusion: mean over the two tokens commutes with the linear out_proj.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


class AudioVisualFusion(nn.Module):
    def __init__(self, audio_dim: int, video_dim: int, output_dim: int, num_attention_heads: int = 8,
                 dropout: float = 0.3):
        super().__init__()
        self.audio_dim, self.video_dim, self.output_dim = audio_dim, video_dim, output_dim
        self.dropout = dropout
        self.audio_projection = nn.Linear(audio_dim, output_dim)
        self.video_projection = nn.Linear(video_dim, output_dim)
        self.cross_attention = nn.MultiheadAttention(embed_dim=output_dim, num_heads=num_attention_heads,
                                                     dropout=dropout, batch_first=True)
        self.fusion_layers = nn.Sequential(nn.Linear(output_dim * 2, output_dim), nn.ReLU(), nn.Dropout(dropout),
                                           nn.LayerNorm(output_dim))

    def forward(self, audio_features, video_features) -> Dict[str, torch.Tensor]:
        E = self.output_dim
        ap = ops.linear(audio_features, self.audio_projection.weight, self.audio_projection.bias)
        vp = ops.linear(video_features, self.video_projection.weight, self.video_projection.bias)
        ca = self.cross_attention
        wv, bv = ca.in_proj_weight[2 * E:], ca.in_proj_bias[2 * E:]
        # audio attends to video -> value token is video (and vice versa)
        a_att = ops.linear(ops.linear(vp, wv, bv), ca.out_proj.weight, ca.out_proj.bias)
        v_att = ops.linear(ops.linear(ap, wv, bv), ca.out_proj.weight, ca.out_proj.bias)
        fl = self.fusion_layers
        y = ops.linear([a_att, v_att], fl[0].weight, fl[0].bias, "relu", dropout=self.dropout, training=self.training)
        y = ops.layer_norm(y, fl[3].weight, fl[3].bias, fl[3].eps)
        # softmax over ONE key: the attention weights are identically 1 (fusion.py:244-255); a cached constant
        ones = ops.constant(1.0, (audio_features.shape[0], 1), audio_features.device)
        return {"fused_features": y, "attention_weights": {"audio_to_video": ones, "video_to_audio": ones}}


class TrimodalFusion(nn.Module):
    def __init__(self, audiovisual_dim: int, text_dim: int, output_dim: int, num_attention_heads: int = 8,
                 dropout: float = 0.3):
        super().__init__()
        self.audiovisual_dim, self.text_dim, self.output_dim = audiovisual_dim, text_dim, output_dim
        self.dropout = dropout
        self.num_heads = num_attention_heads
        self.audiovisual_projection = nn.Linear(audiovisual_dim, output_dim)
        self.text_projection = nn.Linear(text_dim, output_dim)
        self.modality_attention = nn.MultiheadAttention(embed_dim=output_dim, num_heads=num_attention_heads,
                                                        dropout=dropout, batch_first=True)
        self.final_fusion = nn.Sequential(nn.Linear(output_dim, output_dim), nn.ReLU(), nn.Dropout(dropout),
                                          nn.LayerNorm(output_dim))

    def forward(self, audiovisual_features, text_features) -> Dict[str, torch.Tensor]:
        avp = ops.linear(audiovisual_features, self.audiovisual_projection.weight, self.audiovisual_projection.bias)
        tp = ops.linear(text_features, self.text_projection.weight, self.text_projection.bias)
        ma = self.modality_attention
        # packed q|k|v projection of both tokens, written straight into the [B,2,3E] layout
        qkv = ops.grouped_linear([avp, tp], [ma.in_proj_weight] * 2, [ma.in_proj_bias] * 2)
        ctx_mean, attw = ops.mha2_core(qkv, self.num_heads)
        pooled = ops.linear(ctx_mean, ma.out_proj.weight, ma.out_proj.bias)
        ff = self.final_fusion
        y = ops.linear(pooled, ff[0].weight, ff[0].bias, "relu", dropout=self.dropout, training=self.training)
        y = ops.layer_norm(y, ff[3].weight, ff[3].bias, ff[3].eps)
        return {"fused_features": y, "attention_weights": attw}


class UncertaintyAwareGating(nn.Module):
    """Parameter container for fusion.py:346-418.  The reference can never execute it: HierarchicalMultimodalFusion
    passes `uncertainties` positionally to a keyword-only parameter (fusion.py:148-150 vs :384),
    so there is no reference behaviour to reproduce; the parameters exist for state_dict parity."""

    def __init__(self, modality_dims: List[int], output_dim: int, hidden_dim: int = 128):
        super().__init__()
        self.modality_dims = modality_dims
        self.num_modalities = len(modality_dims)
        self.modality_encoders = nn.ModuleList([
            nn.Sequential(nn.Linear(d, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim // 2))
            for d in modality_dims])
        self.uncertainty_encoder = nn.Sequential(nn.Linear(self.num_modalities, hidden_dim // 2), nn.ReLU(),
                                                 nn.Linear(hidden_dim // 2, hidden_dim // 4))
        total = (hidden_dim // 2) * self.num_modalities + hidden_dim // 4
        self.gating_network = nn.Sequential(nn.Linear(total, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, output_dim),
                                            nn.Softmax(dim=-1))

    def forward(self, *modality_features, uncertainties):
        raise NotImplementedError("deer_b200: UncertaintyAwareGating.forward is unreachable in the reference "
                                  "(TypeError at fusion.py:148); no oracle exists for it")


class HierarchicalMultimodalFusion(nn.Module):
    def __init__(self, audio_dim: int, video_dim: int, text_dim: int, fusion_dim: int = 512,
                 intermediate_dim: int = 256, num_attention_heads: int = 8, dropout: float = 0.3,
                 use_uncertainty_weighting: bool = True):
        super().__init__()
        self.audio_dim, self.video_dim, self.text_dim = audio_dim, video_dim, text_dim
        self.fusion_dim, self.intermediate_dim = fusion_dim, intermediate_dim
        self.use_uncertainty_weighting = use_uncertainty_weighting
        self.dropout = dropout
        self.audio_visual_fusion = AudioVisualFusion(audio_dim, video_dim, intermediate_dim, num_attention_heads, dropout)
        self.trimodal_fusion = TrimodalFusion(intermediate_dim, text_dim, fusion_dim, num_attention_heads, dropout)
        if use_uncertainty_weighting:
            self.uncertainty_gate = UncertaintyAwareGating([audio_dim, video_dim, text_dim], output_dim=3)
        self.output_projection = nn.Sequential(nn.Linear(fusion_dim, fusion_dim), nn.ReLU(), nn.Dropout(dropout),
                                               nn.LayerNorm(fusion_dim))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, audio_features, video_features, text_features,
                uncertainties: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if self.use_uncertainty_weighting and uncertainties is not None:
            raise NotImplementedError("deer_b200: the uncertainties= branch raises TypeError in the reference "
                                      "(fusion.py:148-150); only uncertainties=None is defined")
        av = self.audio_visual_fusion(audio_features, video_features)
        tri = self.trimodal_fusion(av["fused_features"], text_features)
        op = self.output_projection
        y = ops.linear(tri["fused_features"], op[0].weight, op[0].bias, "relu", dropout=self.dropout,
                       training=self.training)
        y = ops.layer_norm(y, op[3].weight, op[3].bias, op[3].eps)
        return {"fused_features": y, "audiovisual_features": av["fused_features"],
                "trimodal_features": tri["fused_features"], "av_attention_weights": av["attention_weights"],
                "trimodal_attention_weights": tri["attention_weights"], "uncertainty_weights": None}


def create_fusion_module(fusion_type: str = "hierarchical", **kwargs) -> nn.Module:
    """fusion.py:557 counterpart; only the hierarchical module is on the hot path."""
    if fusion_type != "hierarchical":
        raise NotImplementedError(f"deer_b200: fusion type {fusion_type!r} is never instantiated by the reference "
                                  "callers and is out of scope (SURVEY.md section 2 row 3)")
    return HierarchicalMultimodalFusion(**kwargs)
