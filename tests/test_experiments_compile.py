"""The shelved / head-start kernels under tools/experiments/ are not part of the build; this keeps the ones meant to be
picked up next (unet_conv3x3, onet_chain_fwd) compiling against the current csrc/ headers for sm_100a (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

from .conftest import ROOT

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="no nvcc")
@pytest.mark.parametrize("name", ["unet_conv3x3.cu.txt", "onet_chain_fwd.cu.txt"])
def test_experiment_compiles(name, tmp_path):
    src = os.path.join(ROOT, "tools", "experiments", name)
    r = subprocess.run([NVCC, "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                        "-I", os.path.join(ROOT, "if-defense_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                        "-c", src, "-o", str(tmp_path / "x.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
