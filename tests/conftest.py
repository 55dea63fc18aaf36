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
