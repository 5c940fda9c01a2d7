"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU and exports every symbol that
include/deer_b200.h declares; the host package refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import deer_b200
from deer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "deer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(deer_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/deer_b200.h but not exported"
    # and the Python binding table covers the same set
    assert sorted(_lib.EXPORTS) == syms


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.deer_version() >= 100
    assert isinstance(lib.deer_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    # null pointers / empty shapes are validated on the host before any CUDA call
    rc = lib.deer_gemm(None, 1, 0, None, 1, 0, None, 1, 1, 1, 1, None, 0, 0.0, 1, 0, 0, 0, 0, 0, None)
    assert rc == -1 and b"null" in lib.deer_last_error()
    rc = lib.deer_layernorm_fwd(None, None, None, None, None, None, 0, 0, 1e-5, None)
    assert rc == -1


def test_cpu_tensors_fail_loudly():
    m = deer_b200.MultiDimensionalDEER(64, 3, 32, 0.0)
    with pytest.raises(_lib.DeerError):
        m(torch.randn(2, 64))
    with pytest.raises(_lib.DeerError):
        deer_b200.MultiTaskDEERLoss()({"valence_mu": torch.zeros(2, 1), "valence_nu": torch.ones(2, 1),
                                       "valence_alpha": torch.ones(2, 1) * 2, "valence_beta": torch.ones(2, 1),
                                       "arousal_mu": torch.zeros(2, 1), "arousal_nu": torch.ones(2, 1),
                                       "arousal_alpha": torch.ones(2, 1) * 2, "arousal_beta": torch.ones(2, 1),
                                       "dominance_mu": torch.zeros(2, 1), "dominance_nu": torch.ones(2, 1),
                                       "dominance_alpha": torch.ones(2, 1) * 2, "dominance_beta": torch.ones(2, 1)},
                                      torch.zeros(2, 3))


def test_state_dict_keys_match_reference(golden):
    """Parameter names/shapes are the drop-in contract: the golden fixtures record the
    reference modules' own state_dict shapes."""
    fx = golden("seq_full_b4")
    model = deer_b200.SequenceDEERModel(dropout=0.0)
    sd = model.state_dict()
    for k, shp in fx.shapes.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == tuple(shp), (k, sd[k].shape, shp)
    assert sum(p.numel() for p in model.parameters()) == 9262642


def test_pooled_state_dict_keys_match_reference(golden):
    """CompleteDEERModel (complete_project.py:462): identical parameter names/shapes, 3,918,324 parameters; also
    importable under the `complete_model` name the reference trainer uses (training.py:31)."""
    fx = golden("pooled_b16")
    from deer_b200 import complete_model, complete_project
    assert complete_model.CompleteDEERModel is complete_project.CompleteDEERModel
    model = complete_project.CompleteDEERModel(complete_project.ModelConfig())
    sd = model.state_dict()
    assert set(sd) == set(fx.shapes)
    for k, shp in fx.shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), (k, sd[k].shape, shp)
    assert sum(p.numel() for p in model.parameters()) == 3918324
