import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def conftest_golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def geo():
    return dict(np.load(os.path.join(GOLDEN, "geometry.npz")))


@pytest.fixture(scope="session")
def conv():
    return dict(np.load(os.path.join(GOLDEN, "convonet.npz")))


@pytest.fixture(scope="session")
def conv_sd(conv):
    import torch
    return {k[3:]: torch.from_numpy(v) for k, v in conv.items() if k.startswith("sd/")}


@pytest.fixture(scope="session")
def conv_planes(conv):
    import torch
    return {k: torch.from_numpy(conv["planes_nchw"][i]) for i, k in enumerate(("xz", "xy", "yz"))}


@pytest.fixture(scope="session")
def mathcheck():
    """Host instantiation of the product's __host__ __device__ arithmetic (tests/mathcheck)."""
    import ctypes
    import subprocess
    d = os.path.join(ROOT, "tests", "mathcheck")
    subprocess.check_call(["make", "-C", d, "-s"])
    return ctypes.CDLL(os.path.join(d, "_mathcheck.so"))
