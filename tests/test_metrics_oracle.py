"""CPU: the metrics / linguistic-feature oracle against fixtures produced by the unmodified reference
(src/utils/metrics.py, encoders.py:648-699), and the host-side quantile-edge logic of deer_b200.metrics against
np.quantile."""
import numpy as np
import pytest
import torch

from oracle import deer_oracle as O
from oracle import metrics_oracle as MO

METRIC_FIXTURES = ["metrics_n1000", "metrics_n37", "metrics_nan", "metrics_d1", "metrics_tiny"]


@pytest.mark.parametrize("name", METRIC_FIXTURES)
def test_metrics_oracle_matches_reference(golden, name):
    fx = golden(name)
    a = fx.arrays
    pred, tgt, unc = a["pred"], a["tgt"], a["unc"]
    for i in range(pred.shape[1]):
        assert abs(MO.ccc(tgt[:, i], pred[:, i]) - a["ccc"][i]) <= 2e-6  # the reference reduces in float32
        assert abs(MO.mae(tgt[:, i], pred[:, i]) - a["mae"][i]) <= 1e-6 * abs(a["mae"][i])
        assert abs(MO.rmse(tgt[:, i], pred[:, i]) - a["rmse"][i]) <= 1e-6 * abs(a["rmse"][i])
    with np.errstate(all="ignore"):
        assert abs(MO.uncertainty_calibration_error(pred, tgt, unc) - float(a["uce"])) <= 2e-6
        assert abs(MO.uncertainty_calibration_error(pred, tgt, unc, n_bins=5) - float(a["uce5"])) <= 2e-6
    if "ev_cohens_d" in a:
        for i in range(3):
            assert abs(MO.cohens_d(tgt[:, i], pred[:, i]) - a["ev_cohens_d"][i]) <= 1e-5 * abs(a["ev_cohens_d"][i])
        assert abs(float(a["ev_ece"]) - float(a["uce"])) == 0.0


def test_linguistic_features_wide(golden):
    fx = golden("ling_b32")
    got = O.linguistic_features(torch.from_numpy(fx.arrays["ids"]), torch.from_numpy(fx.arrays["mask"]))
    assert torch.equal(got, torch.from_numpy(fx.arrays["feats"]))


@pytest.mark.parametrize("n,n_bins,seed", [(10, 10, 0), (11, 10, 1), (1003, 10, 2), (4096, 5, 3), (37, 7, 4), (2, 1, 5)])
def test_quantile_edges_host_logic(n, n_bins, seed):
    """deer_b200.metrics rebuilds np.quantile from 2*(n_bins+1) order statistics: bit-identical edges."""
    import deer_b200.metrics as DM
    rng = np.random.default_rng(seed)
    u = np.abs(rng.standard_normal(n)).astype(np.float32)
    if n > 20:
        u[: n // 3] = np.round(u[: n // 3], 1)
    want = np.quantile(u, np.linspace(0, 1, n_bins + 1))
    want[0] = 0
    want[-1] = np.max(u) + 1e-6
    srt = np.sort(u)
    prev, nxt = DM.quantile_ranks(n, n_bins)
    got = DM.quantile_edges(srt[prev], srt[nxt], n, n_bins)
    assert np.array_equal(got, want)
