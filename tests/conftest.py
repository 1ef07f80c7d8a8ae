import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def zipf_inputs():
    return np.load(os.path.join(GOLDEN_DIR, "zipf_inputs.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def aqp():
    """The product binding. Builds libb200aqp.so if it is missing (cross-compile works without a GPU)."""
    import subprocess
    if not os.path.exists(os.path.join(PKG, "libb200aqp.so")):
        subprocess.check_call(["make", "-s", "-C", PKG])
    import b200aqp
    b200aqp.lib()
    return b200aqp


@pytest.fixture(scope="session")
def gpu(aqp):
    aqp.init(0)
    return aqp
