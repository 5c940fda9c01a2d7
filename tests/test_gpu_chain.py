"""The persistent chain kernel (csrc/chain.cu, deer_b200.chain): fusion + NIG head as one launch per direction, against the
module-by-module path (fused 3xTF32 nodes) it replaces -- same outputs, same dropout masks, same parameter and input
gradients -- at several batch sizes (tile tails), with and without dropout, with upstream gradients on the intermediate
outputs, and inside the whole sequence model."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import deer_b200
    from deer_b200 import _lib, chain, ops
    from deer_b200.deer import nig_dict
    from gen_common import det_state_dict, seq_inputs

DEV = "cuda"


@pytest.fixture(autouse=True)
def _policy():
    ops.set_gemm_engine(ops.ENGINE_AUTO)
    ops.set_exact_engine(ops.ENGINE_X3)
    chain.set_enabled(True)
    yield
    chain.set_enabled(False)      # the library default (the chain kernel is an option: measured slower in the full step)


def _modules(dropout, seed=3):
    torch.manual_seed(0)
    fus = deer_b200.HierarchicalMultimodalFusion(512, 512, 512, fusion_dim=512, intermediate_dim=256, dropout=dropout)
    head = deer_b200.MultiDimensionalDEER(512, 3, 256, dropout)
    for m in (fus, head):
        sd = det_state_dict({k: tuple(x.shape) for k, x in m.state_dict().items()}, seed=seed)
        m.load_state_dict({k: x.float() for k, x in sd.items()})
        m.to(DEV).train()
    return fus, head


@pytest.mark.parametrize("B", [4, 33, 96, 256])
@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_chain_matches_module_path(B, dropout):
    g = torch.Generator().manual_seed(B)
    a, v, t = (torch.randn(B, 512, generator=g).to(DEV) for _ in range(3))
    y = torch.tanh(torch.randn(B, 3, generator=g)).to(DEV)
    w_av = torch.randn(B, 256, generator=g).to(DEV) * 1e-3
    w_tri = torch.randn(B, 512, generator=g).to(DEV) * 1e-3
    res = {}
    for use_chain in (False, True):
        fus, head = _modules(dropout)
        ins = [x.clone().requires_grad_(True) for x in (a, v, t)]
        ops.manual_seed(11)
        ops.begin_step()
        before = _lib.launch_count()
        if use_chain:
            assert chain.supported(fus, head, *ins)
            fused, av, tri, attw, ev = chain.fusion_head_chain(fus, head, *ins)
        else:
            f = fus(*ins)
            fused, av, tri, attw = (f["fused_features"], f["audiovisual_features"], f["trimodal_features"],
                                    f["trimodal_attention_weights"])
            ev = head.evidence(fused)
        out = nig_dict(ev, ops.nig_head(ev), head.dimension_names)
        loss = deer_b200.MultiTaskDEERLoss()(out, y)
        # upstream gradients on the intermediate outputs too (they are part of the module's output dictionary)
        total = loss["total_loss"] + (av * w_av).sum() + (tri * w_tri).sum() + fused.sum() * 1e-4
        total.backward()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - before
        grads = {n: p.grad.clone() for n, p in list(fus.named_parameters()) + list(head.named_parameters())
                 if p.grad is not None}
        res[use_chain] = (dict(fused=fused.detach(), av=av.detach(), tri=tri.detach(), attw=attw.detach(),
                               ev=ev.detach()), grads, [x.grad.clone() for x in ins], launches)
    ref, got = res[False], res[True]
    for k in ref[0]:
        assert rel_l2(got[0][k], ref[0][k]) <= 5e-6, (k, rel_l2(got[0][k], ref[0][k]))
    if dropout > 0:     # identical masks: the same units are zero
        assert torch.equal(got[0]["fused"] == 0, ref[0]["fused"] == 0)
    assert set(got[1]) == set(ref[1])
    for n in ref[1]:
        if float(ref[1][n].abs().max()) == 0.0:
            assert float(got[1][n].abs().max()) == 0.0, n          # Q / K rows of the single-key attention
        else:
            assert rel_l2(got[1][n], ref[1][n]) <= 2e-4, (n, rel_l2(got[1][n], ref[1][n]))
    for gx, gs in zip(got[2], ref[2]):
        assert rel_l2(gx, gs) <= 2e-4
    assert got[3] <= 8, got[3]          # chain fwd + NIG head + 2 loss kernels + chain bwd (+ scratch fills on step 1)
    print(f"B={B} dropout={dropout}: launches chain {got[3]} vs modules {ref[3]}")


def test_sequence_model_with_chain_matches_module_path_and_trainer():
    """The whole sequence model, chain on / off: eager autograd and the trainer's direct-gradient step."""
    from deer_b200.trainer import DEERDataParallelTrainer
    B = 32
    raw = seq_inputs(B, 40, 12, 16, seed=9)
    keys = ("audio_features", "video_features", "text_features", "attention_mask", "linguistic_features", "targets")
    batch = {k: x.float().to(DEV) for k, x in zip(keys, raw)}
    res = {}
    for use_chain in (False, True):
        chain.set_enabled(use_chain)
        torch.manual_seed(0)
        model = deer_b200.SequenceDEERModel(dropout=0.3)
        sd = det_state_dict({k: tuple(x.shape) for k, x in model.state_dict().items()}, seed=9)
        model.load_state_dict({k: (x.float() if x.is_floating_point() else x) for k, x in sd.items()})
        model = model.to(DEV).train()
        tr = DEERDataParallelTrainer(model)
        ops.manual_seed(5)
        l = tr.forward_backward(batch).clone()
        res[use_chain] = (l, tr.flat.grads.clone())
    assert torch.allclose(res[True][0], res[False][0], rtol=2e-5, atol=1e-7)
    assert rel_l2(res[True][1], res[False][1]) <= 2e-4
