"""Deterministic tensors shared by the golden-vector generator and the tests.

Weights and inputs are never stored in the fixtures: both sides rebuild them from
(key name, shape, seed) with a CPU torch.Generator, so a fixture only carries the
reference's OUTPUTS.  Biases / norm scales are deliberately non-trivial (the
reference initialises biases to zero, which would hide bias-path bugs).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def det_uniform(key: str, shape, lo: float, hi: float, seed: int = 0) -> torch.Tensor:
    u = torch.rand(tuple(shape), generator=_gen(key, seed), dtype=torch.float64)
    return u * (hi - lo) + lo


def det_normal(key: str, shape, seed: int = 0) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(key, seed), dtype=torch.float64)


def det_param(key: str, shape, seed: int = 0) -> torch.Tensor:
    """Deterministic value for a state_dict entry, chosen by its name/shape."""
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return det_uniform(key, shape, -0.1, 0.1, seed)
    if leaf == "running_var":
        return det_uniform(key, shape, 0.9, 1.3, seed)
    if leaf == "temperature":
        return det_uniform(key, shape, 0.8, 1.2, seed)
    if len(shape) >= 2:
        fan_out = shape[0]
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        a = (6.0 / (fan_in + fan_out)) ** 0.5
        return det_uniform(key, shape, -a, a, seed)
    if "bias" in leaf:
        return det_uniform(key, shape, -0.1, 0.1, seed)
    # 1-D "weight": LayerNorm / BatchNorm scale
    return det_uniform(key, shape, 0.8, 1.2, seed)


def det_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    out = {}
    for k, shp in shapes.items():
        v = det_param(k, shp, seed)
        out[k] = v if v.dtype == torch.int64 else v.to(dtype)
    return out


def probe(key: str, shape, seed: int = 0) -> torch.Tensor:
    """Fixed random cotangent used to turn a tensor output into a scalar."""
    return det_normal("probe:" + key, shape, seed)


def grad_summary(g: torch.Tensor, key: str, seed: int = 0):
    """(l2 norm, dot with a fixed probe): two numbers that pin a big gradient."""
    g = g.detach().double()
    return float(g.norm()), float((g * probe(key, g.shape, seed)).sum())


def seq_inputs(B: int, Ta: int, Tv: int, Tt: int, Da: int = 84, Dv: int = 256, Dt: int = 768, seed: int = 0,
               ragged: bool = True):
    """Synthetic batch with the distributions of SURVEY.md section 8d."""
    audio = det_normal("in:audio", (B, Ta, Da), seed)
    video = det_normal("in:video", (B, Tv, Dv), seed)
    text = det_normal("in:text", (B, Tt, Dt), seed)
    if ragged:
        lens = (det_uniform("in:lens", (B,), 0.0, 1.0, seed) * (Tt - 1)).floor().long() + 1
        lens[0] = Tt
    else:
        lens = torch.full((B,), Tt, dtype=torch.long)
    mask = (torch.arange(Tt)[None, :] < lens[:, None]).double()
    ling = det_uniform("in:ling", (B, 10), 0.0, 1.0, seed)
    targets = torch.tanh(det_normal("in:tgt", (B, 3), seed) + 0.1 * det_normal("in:tgt2", (B, 3), seed))
    return audio, video, text, mask, ling, targets


def pooled_inputs(B: int, seed: int = 0):
    a = det_normal("in:paudio", (B, 84), seed)
    v = det_normal("in:pvideo", (B, 256), seed)
    t = det_normal("in:ptext", (B, 768), seed)
    y = torch.tanh(det_normal("in:ptgt", (B, 3), seed) + 0.1 * det_normal("in:ptgt2", (B, 3), seed))
    return a, v, t, y


def nig_inputs(B: int, seed: int = 0):
    """Raw evidences [B,3,4] and targets [B,3] for loss fixtures."""
    e = det_normal("in:evid", (B, 3, 4), seed) * 1.5
    y = torch.tanh(det_normal("in:ntgt", (B, 3), seed))
    return e, y
