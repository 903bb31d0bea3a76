"""Generates tests/golden/mesh.npz by running the REFERENCE's marching cubes (compiled from /root/reference by
oracle/Makefile into oracle/_ref/libmcubes_ref.so) on seeded volumes.  Run in the build container:
    make -C oracle && python tests/golden/make_mesh_golden.py
Cases: random normal volumes (ragged shape), a volume of small integers (f1 == f2 midpoints, values equal to the
isovalue), and the padded occupancy of a sphere as Generator3D.extract_mesh pads it (generation.py:172-176)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mcubes_ref  # noqa: E402


def main():
    rng = np.random.default_rng(20260117)
    out = {}
    vols = {
        "normal": (rng.standard_normal((9, 7, 11)), 0.1),
        "ints": (rng.integers(-2, 3, size=(8, 8, 6)).astype(np.float64), 0.0),
    }
    g = np.linspace(-0.55, 0.55, 17)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    occ = ((0.37 - np.sqrt(X ** 2 + 0.8 * Y ** 2 + 1.3 * Z ** 2)) * 12).astype(np.float32).astype(np.float64)
    vols["ellipsoid_padded"] = (np.pad(occ, 1, "constant", constant_values=-1e6), float(np.log(0.2) - np.log(0.8)))
    for name, (vol, iso) in vols.items():
        v, f = mcubes_ref.marching_cubes(vol, iso)
        out[name + "_vol"] = vol
        out[name + "_iso"] = np.float64(iso)
        out[name + "_verts"] = v
        out[name + "_faces"] = f.astype(np.int32)
        print(name, vol.shape, v.shape, f.shape)
    np.savez_compressed(os.path.join(HERE, "mesh.npz"), **out)


if __name__ == "__main__":
    main()
