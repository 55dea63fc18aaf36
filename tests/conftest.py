import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: the gpu-marked tests are skipped, not failed (the driver runs the two
    halves separately with -m "not gpu" / -m gpu; this keeps the unfiltered run green on CPU boxes too)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:       # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def tables():
    t = np.load(os.path.join(GOLD, "tables.npz"))
    return t["sub_scores"], t["np_scores"]


@pytest.fixture(scope="session")
def golden():
    def load(name):
        path = os.path.join(GOLD, name)
        if name.endswith(".gz"):
            with gzip.open(path, "rt") as fh:
                return json.load(fh)
        with open(path) as fh:
            return json.load(fh)
    return load


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
