import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Fixture:
    """A golden .npz written by tests/golden/make_golden.py (reference outputs only)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.arrays = {k: z[k] for k in z.files if not k.startswith("__")}
        self.shapes = {k: tuple(v) for k, v in json.loads(bytes(z["__shapes__"]).decode()).items()}
        self.meta = json.loads(bytes(z["__meta__"]).decode())

    def t(self, key, dtype=torch.float64):
        return torch.from_numpy(np.asarray(self.arrays[key])).to(dtype)

    def keys(self, prefix=""):
        return [k for k in self.arrays if k.startswith(prefix)]

    def state_dict(self, dtype=torch.float64):
        from gen_common import det_state_dict
        return det_state_dict(self.shapes, self.meta["seed"], dtype)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Fixture(name)
        return cache[name]
    return load
