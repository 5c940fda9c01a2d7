import torch

from gen_common import grad_summary


ZERO_GRAD = 1e-12  # reference gradients below this are round-off of an exactly-zero derivative


def rel_l2(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def assert_close(got, ref, tol, what=""):
    assert tuple(got.shape) == tuple(ref.shape), (what, got.shape, ref.shape)
    r = rel_l2(got, ref)
    assert r <= tol, f"{what}: rel l2 {r:.3e} > {tol}"


def load_fixture_weights(module, fx, prefix=""):
    sd = {k[len(prefix):]: v for k, v in fx.state_dict(torch.float32).items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not missing, missing
    return module


def check_param_grads(module, fx, seed, tol, prefix=""):
    """Compare parameter gradients with the fixture: whole tensors (`grad:`) or (norm, probe-dot) summaries."""
    params = dict(module.named_parameters())
    n = 0
    for k in fx.keys("grad:"):
        name = k[5:]
        if not name.startswith(prefix):
            continue
        p = params[name[len(prefix):]]
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        ref = fx.t(k)
        if float(ref.abs().max()) == 0.0:
            assert float(g.abs().max()) == 0.0, name
        elif float(ref.abs().max()) < ZERO_GRAD:
            # structurally zero gradient (bias in front of a softmax / batch-stat BatchNorm): round-off only
            assert float(g.abs().max()) < 1e-5, name
        else:
            assert_close(g, ref, tol, name)
        n += 1
    for k in fx.keys("gsum:"):
        name = k[5:]
        if not name.startswith(prefix):
            continue
        p = params[name[len(prefix):]]
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        nrm, dot = grad_summary(g.cpu(), name, seed)
        rn, rd = (float(v) for v in fx.arrays[k])
        if rn == 0.0:
            assert nrm == 0.0, name
            continue
        if rn < ZERO_GRAD:
            assert nrm < 1e-5, name
            continue
        assert abs(nrm - rn) <= tol * rn, (name, nrm, rn)
        assert abs(dot - rd) <= tol * rn * (g.numel() ** 0.5), (name, dot, rd)
        hk = "ghead:" + name
        if hk in fx.arrays:
            ref = fx.t(hk)
            got = g.detach().flatten()[:64].double().cpu()
            assert float((got - ref).abs().max()) <= tol * max(float(ref.abs().max()), rn / g.numel() ** 0.5) * 4, name
        n += 1
    assert n > 0
