"""Device-side validation metrics: the API of src/utils/metrics.py (`DEERMetrics`, `EvaluationResults`,
`uncertainty_calibration_error`) with the reductions done by libdeer_b200 kernels on CUDA tensors instead of a
NumPy/sklearn round trip.

Per call the host reads back 8 fp64 moments per dimension (CCC / MAE / RMSE / Cohen's d are closed forms of them) and,
for the calibration error, 2*(n_bins+1) order statistics and 3*n_bins bin sums.  There is no CPU fallback: inputs must
be CUDA fp32 tensors.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

NMOM = 8
UCE_MAX_RANKS = 24
UCE_WORKSPACE_BYTES = 65536
DIMENSION_NAMES = ("valence", "arousal", "dominance")


@dataclass
class EvaluationResults:
    """metrics.py:29-49"""
    ccc_valence: float
    ccc_arousal: float
    ccc_dominance: float
    mae_valence: float
    mae_arousal: float
    mae_dominance: float
    ece: float
    statistical_significance: Dict[str, float]
    sample_size: int

    @property
    def ccc_average(self) -> float:
        return (self.ccc_valence + self.ccc_arousal + self.ccc_dominance) / 3

    @property
    def mae_average(self) -> float:
        return (self.mae_valence + self.mae_arousal + self.mae_dominance) / 3


def _req2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DeerError(f"deer_b200.metrics: `{name}` must be a CUDA tensor (no CPU fallback on this path)")
    if t.dim() == 1:
        t = t.reshape(-1, 1)
    return t.to(torch.float32).contiguous()


def moments(pred: torch.Tensor, target: torch.Tensor) -> np.ndarray:
    """fp64 [D,8] = {n, sum t, sum p, sum t^2, sum p^2, sum tp, sum |t-p|, sum (t-p)^2} over the non-NaN pairs."""
    p, t = _req2d(pred, "pred"), _req2d(target, "target")
    if p.shape != t.shape:
        raise _lib.DeerError(f"deer_b200.metrics: shape mismatch {tuple(p.shape)} vs {tuple(t.shape)}")
    N, D = p.shape
    if N == 0:
        return np.zeros((D, NMOM))
    out = torch.empty((D, NMOM), device=p.device, dtype=torch.float64)
    call("deer_metrics_moments", ptr(p), ptr(t), N, D, out.data_ptr())
    return out.cpu().numpy()


def ccc_from_moments(m) -> float:
    """metrics.py:59-103: 2*rho*sd_t*sd_p / (var_t + var_p + (mean_t-mean_p)^2) with population moments; 0.0 for an
    empty / constant input (rho undefined)."""
    n = m[0]
    if n == 0:
        return 0.0
    mt, mp = m[1] / n, m[2] / n
    vt, vp = m[3] / n - mt * mt, m[4] / n - mp * mp
    cov = m[5] / n - mt * mp
    if not (vt > 0.0 and vp > 0.0):
        return 0.0
    den = vt + vp + (mt - mp) ** 2
    return float(2.0 * cov / den) if den != 0 else 0.0


def mae_from_moments(m) -> float:
    return float(m[6] / m[0]) if m[0] > 0 else float("inf")


def rmse_from_moments(m) -> float:
    return float(math.sqrt(m[7] / m[0])) if m[0] > 0 else float("inf")


def cohens_d_from_moments(m) -> float:
    """metrics.py:196-208: mean(|e|) / std(|e|) (population std), 0.0 when the std is zero."""
    if m[0] == 0:
        return 0.0
    mean = m[6] / m[0]
    var = m[7] / m[0] - mean * mean
    return float(mean / math.sqrt(var)) if var > 0 else 0.0


def quantile_edges(prev_vals: np.ndarray, next_vals: np.ndarray, n_valid: int, n_bins: int) -> np.ndarray:
    """np.quantile(u, linspace(0,1,n_bins+1)) (method 'linear') rebuilt from the order statistics at floor(vi) and
    floor(vi)+1, vi = (n-1)*q, with numpy's own interpolation arithmetic (fp32 difference, fp64 lerp, the t>=0.5
    branch), then the reference's edits of the first and last edge (metrics.py:253-255)."""
    q = np.linspace(0, 1, n_bins + 1)
    vi = (n_valid - 1) * q
    gamma = vi - np.floor(vi)
    a, b = prev_vals.astype(np.float32), next_vals.astype(np.float32)
    diff = np.subtract(b, a)
    edges = np.add(a, diff * gamma)
    hi = np.subtract(b, diff * (1 - gamma))
    edges = np.where(gamma >= 0.5, hi, edges).astype(np.float64)
    edges[0] = 0
    edges[-1] = np.float32(b[-1]) + 1e-6
    return edges


def quantile_ranks(n_valid: int, n_bins: int):
    q = np.linspace(0, 1, n_bins + 1)
    vi = (n_valid - 1) * q
    prev = np.floor(vi).astype(np.int64)
    nxt = np.minimum(prev + 1, n_valid - 1)
    return prev, nxt


def uncertainty_calibration_error(predictions: torch.Tensor, targets: torch.Tensor, uncertainties: torch.Tensor,
                                  n_bins: int = 10) -> float:
    """metrics.py:214-279 on CUDA tensors: quantile-binned |mean(1-u) - mean(1-err)| weighted by bin population."""
    p, t, u = _req2d(predictions, "predictions"), _req2d(targets, "targets"), _req2d(uncertainties, "uncertainties")
    N, D = p.shape
    if N == 0:
        return 1.0
    if 2 * (n_bins + 1) > UCE_MAX_RANKS:
        raise _lib.DeerError(f"deer_b200.metrics: n_bins={n_bins} unsupported (<= {UCE_MAX_RANKS // 2 - 1})")
    dev = p.device
    err_m = torch.empty(N, device=dev, dtype=torch.float32)
    keys = torch.empty(N, device=dev, dtype=torch.int32)
    n_valid_t = torch.empty(1, device=dev, dtype=torch.int64)
    call("deer_uce_prepare", ptr(p), ptr(t), ptr(u), N, D, ptr(err_m), keys.data_ptr(), n_valid_t.data_ptr())
    n_valid = int(n_valid_t.item())
    if n_valid < n_bins:
        return 1.0
    prev, nxt = quantile_ranks(n_valid, n_bins)
    ranks = np.concatenate([prev, nxt]).astype(np.int64)
    R = ranks.size
    vals = torch.empty(R, device=dev, dtype=torch.float32)
    ws = torch.empty(UCE_WORKSPACE_BYTES, device=dev, dtype=torch.uint8)
    call("deer_uce_select", keys.data_ptr(), N, ranks.ctypes.data_as(ctypes.c_void_p), R, ptr(vals), ws.data_ptr(),
         UCE_WORKSPACE_BYTES)
    v = vals.cpu().numpy()
    edges = quantile_edges(v[:n_bins + 1], v[n_bins + 1:], n_valid, n_bins)
    edges_d = torch.from_numpy(edges).to(dev)
    sums = torch.empty((3, n_bins), device=dev, dtype=torch.float64)
    call("deer_uce_bins", keys.data_ptr(), ptr(err_m), N, n_bins, edges_d.data_ptr(), sums.data_ptr())
    s = sums.cpu().numpy()
    ece = 0.0
    for i in range(n_bins):
        c = s[0, i]
        if c > 0:
            ece += (c / n_valid) * abs(s[1, i] / c - s[2, i] / c)
    return float(ece)


class DEERMetrics:
    """metrics.py:52-211 on CUDA tensors."""

    def __init__(self):
        self.dimension_names = list(DIMENSION_NAMES)

    def concordance_correlation_coefficient(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
        return ccc_from_moments(moments(y_pred, y_true)[0])

    def mean_absolute_error(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
        return mae_from_moments(moments(y_pred, y_true)[0])

    def root_mean_squared_error(self, y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
        return rmse_from_moments(moments(y_pred, y_true)[0])

    def evaluate_predictions(self, predictions: torch.Tensor, targets: torch.Tensor,
                             uncertainties: Optional[torch.Tensor] = None) -> EvaluationResults:
        """metrics.py:127-188: one moments launch for all dimensions + the calibration error."""
        m = moments(predictions, targets)
        res, sig = {}, {}
        for i, name in enumerate(self.dimension_names):
            if i < m.shape[0]:
                res[f"ccc_{name}"] = ccc_from_moments(m[i])
                res[f"mae_{name}"] = mae_from_moments(m[i])
                sig[f"cohens_d_{name}"] = cohens_d_from_moments(m[i])
            else:
                res[f"ccc_{name}"] = 0.0
                res[f"mae_{name}"] = float("inf")
        ece = uncertainty_calibration_error(predictions, targets, uncertainties) if uncertainties is not None else 0.0
        return EvaluationResults(ccc_valence=res["ccc_valence"], ccc_arousal=res["ccc_arousal"],
                                 ccc_dominance=res["ccc_dominance"], mae_valence=res["mae_valence"],
                                 mae_arousal=res["mae_arousal"], mae_dominance=res["mae_dominance"], ece=ece,
                                 statistical_significance=sig, sample_size=int(predictions.shape[0]))

    def compute_all_metrics(self, y_true: torch.Tensor, y_pred: torch.Tensor,
                            uncertainties: Optional[torch.Tensor] = None) -> Dict[str, float]:
        """Flat dict consumed by the compat trainer / evaluation (val_ccc, per-dimension CCC / MAE / RMSE, ece)."""
        m = moments(y_pred, y_true)
        out: Dict[str, float] = {}
        names = self.dimension_names[:m.shape[0]]
        for i, d in enumerate(names):
            out[f"{d}_ccc"] = ccc_from_moments(m[i])
            out[f"{d}_mae"] = mae_from_moments(m[i])
            out[f"{d}_rmse"] = rmse_from_moments(m[i])
        out["mean_ccc"] = float(np.mean([out[f"{d}_ccc"] for d in names]))
        tot_n = m[:, 0].sum()
        out["mean_mae"] = float(m[:, 6].sum() / tot_n) if tot_n > 0 else float("inf")
        if uncertainties is not None:
            out["ece"] = uncertainty_calibration_error(y_pred, y_true, uncertainties)
            mu = moments(uncertainties, uncertainties)
            out["mean_uncertainty"] = float(mu[:, 2].sum() / mu[:, 0].sum()) if mu[:, 0].sum() > 0 else float("nan")
        return out
