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


def test_header_is_plain_c_and_every_symbol_links(lib_path, tmp_path):
    """include/ifd_b200.h compiles as strict C99, a C program that takes the address of EVERY declared entry point links
    against the library, and tests/c_abi/abi_check.c (argument validation through the C ABI) runs without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(lib_path)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                           os.path.join(inc, "ifd_b200.h")])
    src = tmp_path / "all_symbols.c"
    syms = header_symbols()
    src.write_text('#include <stdio.h>\n#include "ifd_b200.h"\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t t[] = {\n'
                   + "".join("    (fn_t)%s,\n" % s for s in syms)
                   + "  };\n  unsigned i, n = 0;\n  for (i = 0; i < sizeof t / sizeof t[0]; ++i) n += t[i] != 0;\n"
                     '  printf("%u\\n", n);\n  return 0;\n}\n')
    exe = tmp_path / "all_symbols"
    link = ["-I", inc, "-L", libdir, "-lifd_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call([gcc, "-std=c99", "-Wall", str(src), "-o", str(exe)] + link)
    assert int(subprocess.check_output([str(exe)]).decode()) == len(syms)
    exe2 = tmp_path / "abi_check"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", os.path.join(ROOT, "tests", "c_abi", "abi_check.c"),
                           "-o", str(exe2)] + link)
    out = subprocess.run([str(exe2)], capture_output=True, text=True)
    assert out.returncode == 0 and "abi ok" in out.stdout, out.stdout
    # the C example of INTEGRATION.md builds and answers with its usage text
    exe3 = tmp_path / "restore_host"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", os.path.join(ROOT, "examples", "restore_host.c"),
                           "-o", str(exe3)] + link)
    out = subprocess.run([str(exe3)], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr
