"""The fused 3xTF32 GEMM engine of the post-pooling chain (csrc/gemm_x3.cu, deer_gemm_x3) against fp64 torch and against the
exact-fp32 SIMT engine with separate elementwise kernels (the round-1 path): plain GEMMs in all four operand
orientations, batch strides, tails and unaligned shapes (K = 10, N = 4), the gate prologue, the bias-gradient side
output, the dropout epilogue (same Philox stream as deer_dropout), cross-CTA split-K, and the fused autograd nodes."""
import pytest
import torch

from helpers import assert_close, cosine, rel_l2

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import deer_b200
    from deer_b200 import _lib, ops

DEV = "cuda"


@pytest.fixture(autouse=True)
def _policy():
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    ops.set_exact_engine(ops.ENGINE_X3)
    yield
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    ops.set_exact_engine(ops.ENGINE_X3)


def _mk(shape, g, scale=1.0):
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale)


@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (256, 1536, 512), (5, 7, 3), (33, 129, 65), (256, 4, 64),
                                   (300, 128, 10), (4096, 256, 768), (20000, 512, 256), (1, 512, 640)])
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_x3_plain_gemm_is_fp32_grade(M, N, K, ta, tb):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A64 = _mk((K, M) if ta else (M, K), g)
    B64 = _mk((N, K) if tb else (K, N), g, 0.1)
    bias64 = _mk((N,), g)
    A, B, bias = A64.float().to(DEV), B64.float().to(DEV), bias64.float().to(DEV)
    C = torch.full((M, N), 7.0, device=DEV)
    ops.gemm_x3(A, A.stride(0), ta, B, B.stride(0), tb, C, N, M, N, K, bias=bias, act=0)
    A64, B64, bias64 = A.double().cpu(), B.double().cpu(), bias.double().cpu()     # the fp32 values the kernel sees
    ref = (A64.t() if ta else A64) @ (B64.t() if tb else B64) + bias64
    # fp32-grade: the error of an exact-fp32 product is ~1e-7 * sqrt(K); TF32 would sit at ~3e-4
    assert rel_l2(C, ref) <= 3e-6, rel_l2(C, ref)
    # accumulate + activation
    C2 = torch.ones((M, N), device=DEV)
    ops.gemm_x3(A, A.stride(0), ta, B, B.stride(0), tb, C2, N, M, N, K, bias=bias, act=1, beta=1.0)
    assert rel_l2(C2, torch.relu(ref + 1.0)) <= 3e-6


def test_x3_engine_through_deer_gemm_and_batch_strides():
    g = torch.Generator().manual_seed(1)
    G, M, N, K = 3, 256, 128, 256
    X = _mk((M, K), g).float().to(DEV)
    W = _mk((G, N, K), g, 0.1).float().to(DEV)
    b = _mk((G, N), g).float().to(DEV)
    out = torch.empty((M, G, N), device=DEV)
    before = _lib.engine_counts()["tf32x3"]
    ops.gemm(X, K, 0, W, K, 1, out, G * N, M, N, K, bias=b, act=1, batch=G, sA=0, sB=N * K, sC=N, sBias=N,
             engine=ops.ENGINE_X3)
    assert _lib.engine_counts()["tf32x3"] == before + 1
    ref = torch.relu(torch.einsum("mk,gnk->mgn", X.double(), W.double()) + b.double())
    assert rel_l2(out, ref) <= 3e-6


@pytest.mark.parametrize("mode,act", [(1, "relu"), (2, "tanh"), (3, "sigmoid")])
@pytest.mark.parametrize("M", [256, 77])
def test_x3_gate_prologue_and_bias_gradient(mode, act, M):
    """dx = (dy * f'(y)) W and dW += (dy * f'(y))^T x with db = column sums, all from the saved output y."""
    g = torch.Generator().manual_seed(mode * 100 + M)
    N, K = 192, 320
    y64 = _mk((M, N), g)
    y64 = {"relu": torch.relu(y64), "tanh": torch.tanh(y64), "sigmoid": torch.sigmoid(y64)}[act]
    dy64, W64, x64 = _mk((M, N), g), _mk((N, K), g, 0.1), _mk((M, K), g)
    scale = 1.0 / 0.7 if mode == 1 else 1.0
    fp = {1: (y64 > 0).double() * scale, 2: 1 - y64 ** 2, 3: y64 * (1 - y64)}[mode]
    dz64 = dy64 * fp
    y, dy, W, x = (t.float().to(DEV) for t in (y64, dy64, W64, x64))
    dx = torch.empty((M, K), device=DEV)
    ops.gemm_x3(dy, N, 0, W, K, 0, dx, K, M, K, N, gate=y, ldgate=N, gate_mode=mode, gate_scale=scale)
    assert rel_l2(dx, dz64 @ W64) <= 3e-6
    dW = torch.ones((N, K), device=DEV)
    db = torch.ones(N, device=DEV)
    ops.gemm_x3(dy, N, 1, x, K, 0, dW, K, N, K, M, beta=1.0, gate=y, ldgate=N, gate_mode=mode, gate_scale=scale,
                colsum=db)
    assert rel_l2(dW, dz64.t() @ x64 + 1.0) <= 3e-6
    assert rel_l2(db, dz64.sum(0) + 1.0) <= 3e-6


def test_x3_split_k_weight_gradient_long_contraction():
    g = torch.Generator().manual_seed(9)
    M, N, K = 20000, 64, 96        # dW [64, 96] over 20000 rows: few tiles, long contraction -> cross-CTA split-K
    dz64, x64 = _mk((M, N), g), _mk((M, K), g)
    dz, x = dz64.float().to(DEV), x64.float().to(DEV)
    dW = torch.zeros((N, K), device=DEV)
    db = torch.zeros(N, device=DEV)
    ops.gemm_x3(dz, N, 1, x, K, 0, dW, K, N, K, M, beta=1.0, colsum=db)
    assert rel_l2(dW, dz64.t() @ x64) <= 3e-6
    assert rel_l2(db, dz64.sum(0)) <= 3e-6


@pytest.mark.parametrize("M,N", [(256, 512), (37, 100), (64, 6)])
def test_x3_dropout_epilogue_matches_dropout_kernel(M, N):
    """The fused epilogue draws the mask deer_dropout draws for the same (seed, offset, step) over the flat index."""
    g = torch.Generator().manual_seed(M + N)
    K = 128
    x, W = _mk((M, K), g).float().to(DEV), _mk((N, K), g, 0.2).float().to(DEV)
    b = _mk((N,), g).float().to(DEV)
    step = torch.tensor([5], dtype=torch.int64, device=DEV)
    plain = torch.empty((M, N), device=DEV)
    ops.gemm_x3(x, K, 0, W, K, 1, plain, N, M, N, K, bias=b, act=1)
    ref = torch.empty_like(plain)
    _lib.call("deer_dropout", plain.data_ptr(), ref.data_ptr(), plain.numel(), 0.3, 1234, 77, step.data_ptr())
    fused = torch.empty((M, N), device=DEV)
    ops.gemm_x3(x, K, 0, W, K, 1, fused, N, M, N, K, bias=b, act=1, drop=(0.3, 1234, 77, step))
    assert torch.equal(fused, ref)
    frac = float((fused == 0).float().mean())
    assert 0.5 < frac < 0.8          # relu zeros + 30 % dropped


def _chain(model_fn, batch_fn, engine):
    ops.set_exact_engine(engine)
    torch.manual_seed(0)
    model = model_fn()
    batch = batch_fn()
    ops.manual_seed(11)
    out = model(*batch)
    return model, out


@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_fused_chain_nodes_match_unfused_simt_path(dropout):
    """fusion + NIG head (train mode) on the fused 3xTF32 nodes == the exact-fp32 SIMT engine with separate dropout /
    bias_act_bwd kernels: same outputs, same dropout masks, same parameter and input gradients."""
    B = 96
    g = torch.Generator().manual_seed(3)
    a, v, t = (torch.randn(B, 512, generator=g).to(DEV) for _ in range(3))
    y = torch.tanh(torch.randn(B, 3, generator=g)).to(DEV)
    res = {}
    for eng in (ops.ENGINE_SIMT, ops.ENGINE_X3):
        ops.set_exact_engine(eng)
        torch.manual_seed(0)
        fus = deer_b200.HierarchicalMultimodalFusion(512, 512, 512, fusion_dim=512, intermediate_dim=256,
                                                      dropout=dropout).to(DEV).train()
        head = deer_b200.MultiDimensionalDEER(512, 3, 256, dropout).to(DEV).train()
        ins = [x.clone().requires_grad_(True) for x in (a, v, t)]
        ops.manual_seed(11)
        ops.begin_step()
        before = _lib.launch_count()
        f = fus(*ins)
        out = head(f["fused_features"])
        loss = deer_b200.MultiTaskDEERLoss()(out, y)
        loss["total_loss"].backward()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - before
        grads = {n: p.grad.clone() for n, p in list(fus.named_parameters()) + list(head.named_parameters())
                 if p.grad is not None}
        res[eng] = (f["fused_features"].detach().clone(), out["mu_all"].detach().clone(), float(loss["total_loss"]),
                    grads, [x.grad.clone() for x in ins], launches)
    s, x = res[ops.ENGINE_SIMT], res[ops.ENGINE_X3]
    assert rel_l2(x[0], s[0]) <= 2e-5 and rel_l2(x[1], s[1]) <= 2e-5
    assert abs(x[2] - s[2]) <= 2e-5 * abs(s[2])
    assert set(x[3]) == set(s[3])
    for n in s[3]:
        if float(s[3][n].abs().max()) == 0.0:
            assert float(x[3][n].abs().max()) == 0.0, n
        else:
            assert rel_l2(x[3][n], s[3][n]) <= 2e-4, (n, rel_l2(x[3][n], s[3][n]))
    for gx, gs in zip(x[4], s[4]):
        assert rel_l2(gx, gs) <= 2e-4
    # the fused nodes launch fewer kernels (no trainer: unbatched heads; how many per-head layers go out as ONE batched
    # launch depends on whether the allocator happened to space their operands evenly, so the margin varies)
    assert x[5] <= 0.9 * s[5], (x[5], s[5])
    print(f"chain launches: fused {x[5]} vs unfused {s[5]}")


@pytest.mark.parametrize("M,N,K", [(12800, 512, 256), (16384, 384, 768), (1300, 256, 520), (4099, 128, 84)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_split_precision_gemm_is_fp32_grade(M, N, K, act):
    """deer_gemm_h16_split: FP16 hi/lo operand pairs, three tcgen05 passes -- fp32-grade products (a single FP16 / TF32
    pass sits at ~3e-4)."""
    g = torch.Generator().manual_seed(M + N + K)
    x = (_mk((M, K), g) * 1.7).float().to(DEV)
    w = _mk((N, K), g, 0.08).float().to(DEV)
    b = _mk((N,), g).float().to(DEV)
    xh, xl, Kp = ops.cast_split16(x)
    wh, wl, _ = ops.cast_split16(w)
    assert rel_l2(xh.double() + xl.double(), x.double()[:, :K] if Kp == K else torch.nn.functional.pad(x.double(), (0, Kp - K))) <= 3e-7
    y = torch.empty((M, N), device=DEV)
    before = _lib.engine_counts()["h16_split"]
    ops.gemm_split(xh, xl, Kp, wh, wl, Kp, y, N, M, N, K, bias=b, act=act)
    assert _lib.engine_counts()["h16_split"] == before + 1
    ref = x.double() @ w.double().t() + b.double()
    ref = {0: ref, 1: torch.relu(ref), 2: torch.tanh(ref)}[act]
    assert rel_l2(y, ref) <= 1e-5, rel_l2(y, ref)
    # the single-pass FP16 product of the same operands, for scale
    y1 = torch.empty((M, N), device=DEV)
    ops.gemm_h16(xh, Kp, 0, wh, Kp, 1, y1, N, M, N, K, bias=b, act=act)
    assert rel_l2(y1, ref) > 20 * rel_l2(y, ref)


def test_split_precision_conv_window_matches_fp64_conv():
    """Conv1d(k=3) forward on the split-precision sliding-window path (overlapping rows of the hi / lo padded copies)."""
    B, T, C = 300, 50, 512
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, C, generator=g).to(DEV)
    w = (torch.randn(C, C, 3, generator=g) * 0.03).to(DEV)
    b = torch.randn(C, generator=g).to(DEV)
    before = _lib.engine_counts()["h16_split"]
    y = ops.conv1d_k3(x, w, b)
    assert _lib.engine_counts()["h16_split"] == before + 1
    ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=1).transpose(1, 2)
    assert rel_l2(y, ref) <= 1e-5, rel_l2(y, ref)
    ops.set_split_forward(False)
    try:
        y_tf32 = ops.conv1d_k3(x, w, b)
    finally:
        ops.set_split_forward(True)
    assert rel_l2(y_tf32, ref) > 1e-4      # what the TF32 taps gave
