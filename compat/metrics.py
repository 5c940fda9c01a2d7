"""Flat module name imported by run_multimodal_deer.py:81 (`from metrics import DEERMetrics`).  Host-side NumPy
metrics (src/utils/metrics.py:59-103 CCC, :214-279 uncertainty calibration error); not on the hot path."""
from typing import Dict, Optional

import numpy as np

DIMS = ("valence", "arousal", "dominance")


class DEERMetrics:
    @staticmethod
    def concordance_correlation_coefficient(y_true, y_pred) -> float:
        """Lin's CCC = 2 cov / (var_t + var_p + (mean_t - mean_p)^2), population moments (metrics.py:59-103)."""
        t, p = np.asarray(y_true, dtype=np.float64).ravel(), np.asarray(y_pred, dtype=np.float64).ravel()
        if t.size < 2:
            return 0.0
        mt, mp = t.mean(), p.mean()
        cov = ((t - mt) * (p - mp)).mean()
        den = t.var() + p.var() + (mt - mp) ** 2
        return float(2.0 * cov / den) if den > 0 else 0.0

    @staticmethod
    def uncertainty_calibration_error(y_true, y_pred, uncertainties, n_bins: int = 10) -> float:
        """Binned |mean confidence - mean accuracy| with confidence = 1/(1+u), accuracy = 1 - |err| (metrics.py:214-279)."""
        err = np.abs(np.asarray(y_true, dtype=np.float64) - np.asarray(y_pred, dtype=np.float64)).ravel()
        conf = 1.0 / (1.0 + np.asarray(uncertainties, dtype=np.float64).ravel())
        edges = np.linspace(0.0, 1.0, n_bins + 1)
        ece = 0.0
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = (conf > lo) & (conf <= hi)
            if m.any():
                ece += m.mean() * abs(conf[m].mean() - (1.0 - err[m].mean()))
        return float(ece)

    def compute_all_metrics(self, y_true, y_pred, uncertainties: Optional[np.ndarray] = None) -> Dict[str, float]:
        t, p = np.asarray(y_true, dtype=np.float64), np.asarray(y_pred, dtype=np.float64)
        out = {}
        for i, d in enumerate(DIMS[:t.shape[1]]):
            out[f"{d}_ccc"] = self.concordance_correlation_coefficient(t[:, i], p[:, i])
            out[f"{d}_mae"] = float(np.abs(t[:, i] - p[:, i]).mean())
            out[f"{d}_rmse"] = float(np.sqrt(((t[:, i] - p[:, i]) ** 2).mean()))
        out["mean_ccc"] = float(np.mean([out[f"{d}_ccc"] for d in DIMS[:t.shape[1]]]))
        out["mean_mae"] = float(np.abs(t - p).mean())
        if uncertainties is not None:
            out["ece"] = self.uncertainty_calibration_error(t, p, uncertainties)
            out["mean_uncertainty"] = float(np.asarray(uncertainties).mean())
        return out
