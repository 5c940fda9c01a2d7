"""Persistent tcgen05 cluster LSTM (csrc/lstm_persistent.cu) against the exact-fp32 stepwise engine and the oracle."""
import pytest
import torch

from deer_b200 import ops
from helpers import assert_close, cosine
from oracle import deer_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_layer(In, H, seed):
    g = torch.Generator().manual_seed(seed)
    a = (6.0 / (In + 4 * H)) ** 0.5
    b = (6.0 / (H + 4 * H)) ** 0.5
    ws = []
    for _ in range(2):
        ws += [(torch.rand(4 * H, In, generator=g) * 2 - 1) * a, (torch.rand(4 * H, H, generator=g) * 2 - 1) * b,
               (torch.rand(4 * H, generator=g) * 2 - 1) * 0.1, (torch.rand(4 * H, generator=g) * 2 - 1) * 0.1]
    return ws


def run(x, ws, engine, grad):
    ops.set_lstm_engine(engine)
    xs = x.to(DEV).requires_grad_(grad)
    wd = [w.to(DEV).requires_grad_(grad) for w in ws]
    with torch.set_grad_enabled(grad):
        h = ops.bilstm_layer(ops.to_time_major(xs), *wd)
    if grad:
        g = torch.Generator().manual_seed(5)
        pr = torch.randn(h.shape, generator=g).to(DEV)
        (h * pr).sum().backward()
        return h.detach(), [xs.grad] + [w.grad for w in wd]
    return h, None


@pytest.mark.parametrize("B,T,In", [(4, 6, 84), (64, 20, 84), (70, 9, 512), (130, 33, 84)])
@pytest.mark.parametrize("grad", [False, True])
def test_persistent_matches_stepwise(B, T, In, grad):
    H = 256
    ws = make_layer(In, H, B + T)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(B))
    ops.set_gemm_engine(ops.ENGINE_SIMT)     # exact input projection for both runs: isolates the recurrence
    h_ref, g_ref = run(x, ws, ops.ENGINE_SIMT, grad)
    h_per, g_per = run(x, ws, ops.ENGINE_AUTO, grad)
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    assert_close(h_per, h_ref, 1e-3, "h")
    assert float((h_per - h_ref).abs().max()) < 2e-3
    if grad:
        for a, b_ in zip(g_per, g_ref):
            assert cosine(a, b_) > 0.9999


def test_persistent_full_length_vs_oracle():
    B, T, In, H = 8, 300, 84, 256
    ws = make_layer(In, H, 77)
    x = torch.randn(B, T, In, generator=torch.Generator().manual_seed(1))
    h, _ = run(x, ws, ops.ENGINE_AUTO, False)
    wd = [w.double() for w in ws]
    f = O.lstm_direction(x.double(), *wd[:4], False)
    r = O.lstm_direction(x.double(), *wd[4:], True)
    ref = torch.cat([f, r], dim=-1).permute(1, 0, 2)
    assert_close(h, ref, 1e-3, "h vs oracle")
