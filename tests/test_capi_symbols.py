"""The C-ABI library loads and exports every symbol include/ifd_b200.h declares (no compute without a GPU),
and the ctypes signature table covers exactly that set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ifd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ifd_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib_path():
    from ifdefense_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.LIB_PATH


def test_exports_match_header(lib_path):
    from ifdefense_b200 import capi
    L = ctypes.CDLL(lib_path)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s
    assert sorted(capi.SIGNATURES) == syms


def test_abi_version_and_defaults(lib_path):
    from ifdefense_b200 import capi
    L = capi.lib()
    assert L.ifd_abi_version() == 1
    p = capi.default_params()
    assert (p.n_steps, p.knn_k, p.normalize_out) == (201, 5, 1)
    assert (p.lr, p.beta1, p.beta2, p.adam_eps) == (1e-3, 0.9, 0.999, 1e-8)
    assert (p.occ_target, p.rep_weight, p.rep_radius, p.rep_h, p.rep_eps, p.padding) == (0.2, 500.0, 0.07, 0.03, 1e-12, 0.1)
    assert L.ifd_convonet_decoder_nfloats(32, 32, 5) == 16001
    assert L.ifd_convonet_opt_workspace_bytes(64, 1024) > 64 * 1024 * 3 * 8 * 6


def test_argument_errors_without_gpu(lib_path):
    """Validation happens before any CUDA call, and the message travels through ifd_last_error()."""
    from ifdefense_b200 import capi
    L = capi.lib()
    assert L.ifd_knn(None, 1, 8, 3, 2, 1, None, None, None) == -1
    assert b"null" in L.ifd_last_error()
    with pytest.raises(RuntimeError):
        capi.check(L.ifd_fps(None, 1, 8, 2, None, None, None), "ifd_fps")
    with pytest.raises(RuntimeError):
        capi.default_params(no_such_field=1)


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ifdefense_b200 import defense
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        defense.knn_point(5, torch.zeros(1, 16, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        defense.repulsion_loss(torch.zeros(1, 16, 3))
