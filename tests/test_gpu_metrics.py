"""GPU parity of the section-8f kernels (linguistic features, validation metrics) through the C ABI: bit-exact for the
integer / index work (feature rows, order statistics, bin populations), fp64-accumulated sums within 2e-6 of the
reference's float32 NumPy reductions and 1e-10 of the fp64 oracle."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import deer_oracle as O
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def _dm():
    import deer_b200.metrics as DM
    return DM


@pytest.mark.parametrize("name", ["ling", "ling_b32"])
def test_linguistic_features_golden(golden, name):
    from deer_b200 import ops
    fx = golden(name)
    ids, mask = torch.from_numpy(fx.arrays["ids"]).cuda(), torch.from_numpy(fx.arrays["mask"]).cuda()
    got = ops.linguistic_features(ids, mask).cpu()
    assert torch.equal(got, torch.from_numpy(fx.arrays["feats"]))


@pytest.mark.parametrize("B,T,vocab", [(1024, 64, 30522), (64, 512, 50), (3, 1, 5), (257, 33, 1100)])
def test_linguistic_features_oracle(B, T, vocab):
    from deer_b200 import ops
    g = torch.Generator().manual_seed(B + T)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    mask = (torch.rand((B, T), generator=g) < 0.8).long()
    mask[0] = 0
    want = O.linguistic_features(ids, mask)
    got = ops.linguistic_features(ids.cuda(), mask.cuda()).cpu()
    assert torch.equal(got, want)
    # float / bool masks select the same tokens
    got2 = ops.linguistic_features(ids.cuda(), mask.float().cuda()).cpu()
    assert torch.equal(got2, want)


def test_text_encoder_computes_features_from_ids():
    import deer_b200
    torch.manual_seed(0)
    enc = deer_b200.EnhancedTextEncoder({"dropout": 0.0}).cuda().eval()
    B, T = 8, 16
    emb = torch.randn(B, T, 768, device="cuda")
    ids = torch.randint(0, 2000, (B, T), device="cuda")
    mask = torch.ones(B, T, device="cuda")
    ling = enc.extract_linguistic_features(ids, mask)
    a = enc(emb, mask, input_ids=ids)
    b = enc(emb, mask, ling)
    assert torch.equal(a, b)
    assert not torch.equal(a, enc(emb, mask))  # zeros when neither is given


METRIC_FIXTURES = ["metrics_n1000", "metrics_n37", "metrics_nan", "metrics_d1", "metrics_tiny"]


@pytest.mark.parametrize("name", METRIC_FIXTURES)
def test_metrics_golden(golden, name):
    DM = _dm()
    a = golden(name).arrays
    pred, tgt, unc = (torch.from_numpy(a[k]).cuda() for k in ("pred", "tgt", "unc"))
    m = DM.moments(pred, tgt)
    for i in range(pred.shape[1]):
        assert abs(DM.ccc_from_moments(m[i]) - a["ccc"][i]) <= 2e-6
        assert abs(DM.mae_from_moments(m[i]) - a["mae"][i]) <= 2e-6 * abs(a["mae"][i])
        assert abs(DM.rmse_from_moments(m[i]) - a["rmse"][i]) <= 2e-6 * abs(a["rmse"][i])
        # fp64 oracle on the same float32 inputs
        assert abs(DM.ccc_from_moments(m[i]) - MO.ccc(a["tgt"][:, i], a["pred"][:, i])) <= 1e-10
    assert abs(DM.uncertainty_calibration_error(pred, tgt, unc) - float(a["uce"])) <= 2e-6
    assert abs(DM.uncertainty_calibration_error(pred, tgt, unc, n_bins=5) - float(a["uce5"])) <= 2e-6
    with np.errstate(all="ignore"):
        assert abs(DM.uncertainty_calibration_error(pred, tgt, unc) -
                   MO.uncertainty_calibration_error(a["pred"], a["tgt"], a["unc"])) <= 1e-10
    if "ev_ccc" in a:
        ev = DM.DEERMetrics().evaluate_predictions(pred, tgt, unc)
        got = np.array([ev.ccc_valence, ev.ccc_arousal, ev.ccc_dominance])
        assert np.abs(got - a["ev_ccc"]).max() <= 2e-6
        got = np.array([ev.mae_valence, ev.mae_arousal, ev.mae_dominance])
        assert np.abs(got - a["ev_mae"]).max() <= 2e-6
        assert abs(ev.ece - float(a["ev_ece"])) <= 2e-6
        d = np.array([ev.statistical_significance[f"cohens_d_{k}"] for k in ("valence", "arousal", "dominance")])
        assert np.abs(d / a["ev_cohens_d"] - 1).max() <= 1e-5
        assert ev.sample_size == pred.shape[0]


def test_radix_select_is_exact():
    """Order statistics through the C ABI: bit-exact against a full sort, with negatives, ties, zeros of both signs'
    neighbours and dropped (NaN / inf) samples."""
    from deer_b200 import _lib
    DM = _dm()
    N = 200_003
    g = torch.Generator().manual_seed(3)
    u = torch.randn(N, generator=g)
    u[:5000] = torch.round(u[:5000] * 4) / 4
    u[7] = float("nan")
    u[9] = float("inf")
    u[11] = float("-inf")
    ud = u.cuda().reshape(N, 1)
    z = torch.zeros(N, 1, device="cuda")
    err = torch.empty(N, device="cuda")
    keys = torch.empty(N, device="cuda", dtype=torch.int32)
    nv = torch.empty(1, device="cuda", dtype=torch.int64)
    _lib.call("deer_uce_prepare", z.data_ptr(), z.data_ptr(), ud.data_ptr(), N, 1, err.data_ptr(), keys.data_ptr(),
              nv.data_ptr())
    n_valid = int(nv.item())
    keep = ~(torch.isnan(u) | torch.isinf(u))
    assert n_valid == int(keep.sum()) == N - 3
    srt = torch.sort(u[keep]).values
    ranks = np.array([0, 1, 2, 4999, 5000, n_valid // 2, n_valid // 2 + 1, n_valid - 2, n_valid - 1, 12345, 77777],
                     dtype=np.int64)
    vals = torch.empty(ranks.size, device="cuda")
    ws = torch.empty(DM.UCE_WORKSPACE_BYTES, device="cuda", dtype=torch.uint8)
    _lib.call("deer_uce_select", keys.data_ptr(), N, ranks.ctypes.data_as(ctypes.c_void_p), ranks.size,
              vals.data_ptr(), ws.data_ptr(), DM.UCE_WORKSPACE_BYTES)
    got = vals.cpu()
    want = srt[torch.from_numpy(ranks)]
    assert torch.equal(got, want)  # value equality: the sort may order -0.0 / +0.0 either way


def test_metrics_large_vs_oracle():
    DM = _dm()
    N = 1 << 20
    g = torch.Generator().manual_seed(5)
    tgt = torch.tanh(torch.randn(N, 3, generator=g))
    pred = 0.6 * tgt + 0.4 * torch.randn(N, 3, generator=g)
    unc = (0.5 * torch.randn(N, 3, generator=g)).abs() + 0.01
    m = DM.moments(pred.cuda(), tgt.cuda())
    for i in range(3):
        assert abs(DM.ccc_from_moments(m[i]) - MO.ccc(tgt[:, i].numpy(), pred[:, i].numpy())) <= 1e-9
        assert abs(DM.mae_from_moments(m[i]) - MO.mae(tgt[:, i].numpy(), pred[:, i].numpy())) <= 1e-9
    got = DM.uncertainty_calibration_error(pred.cuda(), tgt.cuda(), unc.cuda())
    want = MO.uncertainty_calibration_error(pred.numpy(), tgt.numpy(), unc.numpy())
    assert abs(got - want) <= 1e-9


def test_metrics_edge_cases():
    DM = _dm()
    dm = DM.DEERMetrics()
    c = torch.ones(16, device="cuda")
    assert dm.concordance_correlation_coefficient(c, c) == 0.0  # constant input: correlation undefined -> 0.0
    e = torch.empty(0, 3, device="cuda")
    assert DM.uncertainty_calibration_error(e, e, e) == 1.0
    s = torch.rand(5, 3, device="cuda")
    assert DM.uncertainty_calibration_error(s, s, s) == 1.0  # fewer kept samples than bins (metrics.py:246-247)
    with pytest.raises(Exception):
        DM.moments(torch.ones(4, 3), torch.ones(4, 3))  # CPU tensors: no fallback
