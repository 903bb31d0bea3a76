"""Generates tests/golden/mise.npz by running the REFERENCE's MISE: ONet/im2mesh/utils/libmise/mise.pyx is cythonized
from a scratch copy under /tmp (the reference tree is read-only; nothing of it enters the repository) and driven by the
loop of Generator3D.generate_from_latent (generation.py:113-130) on the reproducible fields of oracle/mise_port.py.
Stored per case: the points queried per round, the dense result (small lattices) or its digest (129^3).
    python tests/golden/make_mise_golden.py"""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mise_port  # noqa: E402

REF = "/root/reference/ONet/im2mesh/utils/libmise/mise.pyx"


def build_reference_mise():
    d = tempfile.mkdtemp(prefix="ref_mise_")
    shutil.copy(REF, d)
    with open(os.path.join(d, "setup.py"), "w") as f:
        f.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\n"
                "setup(ext_modules=cythonize([Extension('mise', ['mise.pyx'], language='c++')], language_level=3))\n")
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=d, stdout=subprocess.DEVNULL)
    sys.path.insert(0, d)
    import mise
    return mise.MISE


def run_reference(MISE, eval_fn, r0, depth, thr):
    ex = MISE(r0, depth, thr)
    rounds = []
    pts = ex.query()
    while pts.shape[0] != 0:
        rounds.append(pts.shape[0])
        ex.update(pts, np.asarray(eval_fn(pts), dtype=np.float64))
        pts = ex.query()
    return ex.to_dense(), rounds


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    MISE = build_reference_mise()
    out = {}
    cases = [("s0", 8, 2, 0, 0.5), ("s1", 8, 2, 1, -1.38629), ("s2", 6, 1, 2, 0.0), ("s3", 4, 3, 1, 0.2), ("d0", 5, 0, 0, 0.0),
             ("big0", 32, 2, 0, -1.38629), ("big1", 32, 2, 1, -1.38629)]
    for name, r0, depth, kind, thr in cases:
        res = r0 << depth
        fn = lambda p, res=res, kind=kind: mise_port.analytic_field(p, res, kind)
        dense, rounds = run_reference(MISE, fn, r0, depth, thr)
        mine, my_rounds = mise_port.mise_loop(fn, r0, depth, thr)
        assert rounds == my_rounds and np.array_equal(dense, mine), name          # the restatement, checked on the spot
        out[name + "_cfg"] = np.array([r0, depth, kind], dtype=np.int64)
        out[name + "_thr"] = np.float64(thr)
        out[name + "_rounds"] = np.array(rounds, dtype=np.int64)
        out[name + "_sha256"] = np.array(digest(dense))
        if res <= 32:
            out[name + "_dense"] = dense.astype(np.float32) if np.array_equal(dense.astype(np.float32), dense) else dense
        print(name, "lattice", res + 1, "rounds", rounds, "evaluated", sum(rounds), "of", (res + 1) ** 3)
    np.savez_compressed(os.path.join(HERE, "mise.npz"), **out)


if __name__ == "__main__":
    main()
