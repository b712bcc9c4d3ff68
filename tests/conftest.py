import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        raw = json.load(f)

    def dec(d):
        if isinstance(d, dict) and "shape" in d:
            a = np.array(d["re"], dtype=float)
            if "im" in d:
                a = a + 1j * np.array(d["im"], dtype=float)
            return a.reshape(d["shape"])
        return d
    return {k: dec(v) for k, v in raw.items()}


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def qlib():
    """The product library (built in-tree if missing); host-only entry points work without a GPU."""
    from qinchworm_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    lib.load()
    return lib


@pytest.fixture(scope="session")
def gpu_ctx(qlib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return qlib.Context()
