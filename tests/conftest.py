import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "maze-solving-agent-gymnasium_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"])) if "meta" in z.files else None
    return z, meta


@pytest.fixture(scope="session")
def golden_steps():
    return load_golden("steps")


@pytest.fixture(scope="session")
def golden_bestdir():
    return load_golden("bestdir")


@pytest.fixture(scope="session")
def golden_metrics():
    return load_golden("metrics")


@pytest.fixture(scope="session")
def golden_qagent():
    return load_golden("qagent")
