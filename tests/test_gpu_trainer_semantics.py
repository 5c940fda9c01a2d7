"""Trainer-step semantics against the reference trainer (src/training/training.py) and stock torch.optim.AdamW:
dataset loss weighting (:59-61, :211-212), parameter groups and CosineAnnealingLR per group (:121-159), parameters the
reference optimizer never touches, the objective read from model.loss_fn, dropout masks per step, the device stager."""
import copy
import math

import pytest
import torch

from helpers import assert_close, cosine

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import deer_b200
    from deer_b200 import ops
    from deer_b200.data import DevicePrefetcher
    from deer_b200.trainer import GROUP_DEFAULT, GROUP_ENCODER, GROUP_FROZEN, DEERDataParallelTrainer
    from gen_common import det_state_dict, pooled_inputs
    from oracle import deer_oracle as O

DEV = "cuda"


def _pooled(seed=61, dropout=0.0):
    torch.manual_seed(0)
    model = deer_b200.CompleteDEERModel(deer_b200.ModelConfig(dropout=dropout))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd64 = det_state_dict(shapes, seed=seed)
    model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in sd64.items()})
    model = model.to(DEV).train()
    if dropout == 0.0:
        for mod in model.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    return model, sd64


def _batch(B, seed):
    a, v, t, y = pooled_inputs(B, seed)
    return {"audio_features": a.float().to(DEV), "video_features": v.float().to(DEV),
            "text_features": t.float().to(DEV), "targets": y.float().to(DEV),
            "dataset_id": torch.ones(B, 1, dtype=torch.int64, device=DEV)}, (a, v, t, y)


def test_dataset_loss_weight_matches_reference_weighted_backward():
    """training.py:211-212: `(total_loss * dataset_weight).backward()` -- gradients scale with the weight, the reported
    loss components do not; checked against the fp64 oracle's weighted backward."""
    model, sd64 = _pooled()
    tr = DEERDataParallelTrainer(model)
    batch, (a, v, t, y) = _batch(48, 61)
    l1 = tr.forward_backward(batch, loss_weight=1.0).clone()
    g1 = tr.flat.grads.clone()
    l2 = tr.forward_backward(batch, loss_weight=0.6).clone()
    g2 = tr.flat.grads.clone()
    assert torch.allclose(l1, l2, rtol=1e-6, atol=0)
    assert_close(g2, 0.6 * g1, 1e-5, "weighted gradient")
    sdg = {k: (w.clone().requires_grad_(True) if w.is_floating_point() else w) for k, w in sd64.items()}
    ref = O.pooled_model(a, v, t, sdg)
    rl = O.multitask_deer_loss(O.pooled_loss_inputs(ref), y)
    (rl["total_loss"] * 0.6).backward()
    got, want = [], []
    for n, o in zip(tr.flat.names, tr.flat.offsets):
        og = sdg[n].grad
        if og is None:
            continue
        got.append(g2[o:o + og.numel()].double().cpu())
        want.append(og.flatten())
    got, want = torch.cat(got), torch.cat(want)
    assert cosine(got, want) >= 0.9999
    assert abs(float(got.norm() / want.norm()) - 1.0) <= 1e-3


def test_trainer_step_matches_torch_adamw_with_reference_groups():
    """One fused clip + AdamW step == torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW with the reference's groups:
    names containing 'encoder' at 0.5 x lr, weight decay on everything that has a gradient, and parameters whose
    reference gradient is None (calibration layer) untouched."""
    model, _ = _pooled(seed=62)
    shadow = copy.deepcopy(model)
    lr, wd = 1e-3, 1e-2
    tr = DEERDataParallelTrainer(model, learning_rate=lr, weight_decay=wd, gradient_clip=0.5)
    assert [g for g, _, _ in tr.flat.group_bounds] == [GROUP_ENCODER, GROUP_DEFAULT, GROUP_FROZEN]
    batch, _ = _batch(32, 62)
    frozen_before = {n: p.detach().clone() for n, p in model.named_parameters() if n.startswith("calibration_layer.")}
    assert frozen_before
    tr.set_group_lrs({GROUP_ENCODER: 0.5 * lr, GROUP_DEFAULT: lr})
    tr.forward_backward(batch)
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    tr.optimizer_step()
    torch.cuda.synchronize()
    # stock optimizer on the shadow copy with the SAME gradients (fp64 to separate algorithm from rounding)
    shadow = shadow.double()
    named = dict(shadow.named_parameters())
    enc = [p for n, p in named.items() if "encoder" in n and not n.startswith("calibration_layer.")]
    rest = [p for n, p in named.items() if "encoder" not in n and not n.startswith("calibration_layer.")]
    opt = torch.optim.AdamW([{"params": enc, "lr": 0.5 * lr}, {"params": rest, "lr": lr}], weight_decay=wd, eps=1e-8)
    for n, p in named.items():
        if not n.startswith("calibration_layer."):
            p.grad = grads[n].double()
    torch.nn.utils.clip_grad_norm_(enc + rest, 0.5)
    opt.step()
    for n, p in model.named_parameters():
        if n.startswith("calibration_layer."):
            assert torch.equal(p.detach(), frozen_before[n]), n       # no decay, no update
        else:
            d = float((p.detach().double() - named[n].detach()).abs().max())
            assert d <= 2e-6 * max(1.0, float(named[n].abs().max())), (n, d)


def test_group_learning_rates_follow_cosine_annealing_per_group():
    """compat DEERTrainer schedule == torch CosineAnnealingLR(T_max, eta_min=1e-6) on both reference groups."""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))
    training = importlib.import_module("training")
    model, _ = _pooled(seed=63)
    cfg = training.TrainingConfig(learning_rate=2e-4, num_epochs=10)
    t = training.DEERTrainer(model, cfg, torch.device(DEV))
    p0, p1 = torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([{"params": [p0], "lr": 1e-4}, {"params": [p1], "lr": 2e-4}])
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10, eta_min=1e-6)
    for epoch in range(10):
        want = [g["lr"] for g in opt.param_groups]
        got = t._group_lrs(epoch)
        assert math.isclose(got[GROUP_ENCODER], want[0], rel_tol=1e-9, abs_tol=1e-15)
        assert math.isclose(got[GROUP_DEFAULT], want[1], rel_tol=1e-9, abs_tol=1e-15)
        opt.step()
        sched.step()
    t.step.set_group_lrs(t._group_lrs(3))
    dev = t.step.group_lrs()
    assert math.isclose(dev[GROUP_ENCODER], t._group_lrs(3)[GROUP_ENCODER], rel_tol=1e-6)


def test_trainer_objective_follows_model_loss_fn():
    """The optimised objective is model.loss_fn's: non-default reg / kl / ece / cross-dimension / task weights reach the
    fused loss kernel (and equal what compute_loss reports)."""
    model, _ = _pooled(seed=64)
    model.loss_fn = deer_b200.MultiTaskDEERLoss(task_weights={"valence": 2.0, "arousal": 1.0, "dominance": 0.5},
                                                cross_dim_weight=0.2, reg_weight=0.3, kl_weight=0.05, ece_weight=0.0)
    tr = DEERDataParallelTrainer(model)
    batch, _ = _batch(40, 64)
    losses = tr.forward_backward(batch)
    g_tr = tr.flat.grads.clone()
    model.zero_grad()
    out = model(batch)
    rep = model.compute_loss(out, batch["targets"])
    assert abs(float(losses[-1]) - float(rep["total_loss"])) <= 1e-6 * abs(float(rep["total_loss"]))
    for p in model.parameters():
        p.grad = None
    rep["total_loss"].backward()
    got, want = [], []
    for n, p in model.named_parameters():
        if p.grad is not None:
            o = tr.flat.offsets[tr.flat.names.index(n)]
            got.append(g_tr[o:o + p.numel()])
            want.append(p.grad.flatten())
    assert_close(torch.cat(got), torch.cat(want), 1e-4, "trainer gradient vs autograd of compute_loss")


def test_dropout_masks_change_between_forwards_without_a_trainer():
    """A plain torch.optim loop around the drop-in modules must not train one fixed sub-network: two consecutive
    train-mode forwards draw different masks; a trainer's step tensor is bound only during its own step."""
    torch.manual_seed(0)
    model = deer_b200.CompleteDEERModel(deer_b200.ModelConfig(dropout=0.3)).to(DEV).train()
    batch, _ = _batch(64, 65)
    o1 = model(batch)["fused_features"].clone()
    o2 = model(batch)["fused_features"].clone()
    assert float((o1 - o2).abs().max()) > 1e-3
    x = torch.ones(1 << 14, device=DEV)
    ops.begin_step()
    a = ops.dropout(x, 0.5, True)
    ops.begin_step()
    b = ops.dropout(x, 0.5, True)
    assert 0.3 < float((a != b).float().mean()) < 0.7
    tr1 = DEERDataParallelTrainer(model)
    assert ops._dropout_state["step"] is None        # constructing a trainer binds nothing globally
    tr1.forward_backward(batch)
    assert ops._dropout_state["step"] is None
    model.eval()
    e1 = model(batch)["fused_features"].clone()
    e2 = model(batch)["fused_features"].clone()
    assert torch.equal(e1, e2)


def test_device_prefetcher_stages_dicts_and_tuples_into_static_buffers():
    g = torch.Generator().manual_seed(5)
    host = []
    for i in range(5):
        B = 8 if i < 4 else 3           # short last batch: its own buffer sets
        host.append((torch.randn(B, 84, generator=g), torch.randn(B, 256, generator=g),
                     torch.randn(B, 768, generator=g), torch.randn(B, 3, generator=g)))
    st = DevicePrefetcher(host, DEV, depth=2)
    seen, ptrs = [], []
    for dev in st:
        assert set(dev) == {"audio_features", "video_features", "text_features", "targets"}
        seen.append({k: v.clone() for k, v in dev.items()})
        ptrs.append(dev["audio_features"].data_ptr())
    torch.cuda.synchronize()
    assert len(seen) == 5 and ptrs[0] == ptrs[2] and ptrs[1] == ptrs[3] and ptrs[0] != ptrs[1]
    for h, d in zip(host, seen):
        for x, k in zip(h, ("audio_features", "video_features", "text_features", "targets")):
            assert torch.equal(d[k].cpu(), x)
    assert st.bytes_per_batch == 3 * (84 + 256 + 768 + 3) * 4
    dict_batches = [{"audio_features": h[0], "video_features": h[1], "text_features": h[2], "targets": h[3],
                     "dataset_id": torch.zeros(h[0].shape[0], 1, dtype=torch.int64)} for h in host[:2]]
    out = list(DevicePrefetcher(dict_batches, DEV))
    assert out[0]["dataset_id"].dtype == torch.int64 and out[0]["audio_features"].is_cuda
