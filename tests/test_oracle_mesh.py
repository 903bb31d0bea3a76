"""CPU: the oracle's marching-cubes restatement against the reference-generated fixture (tests/golden/mesh.npz), against
the reference's own compiled code where oracle/_ref exists, and the generated case tables' consistency."""
import os
import re

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import mcubes_ref, mise_port

from .conftest import GOLDEN, ROOT

CASES = ["normal", "ints", "ellipsoid_padded"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "mesh.npz"))


@pytest.mark.parametrize("name", CASES)
def test_restatement_equals_reference_fixture(gold, name):
    v, f = co.marching_cubes(gold[name + "_vol"], float(gold[name + "_iso"]))
    assert np.array_equal(v, gold[name + "_verts"])          # bit for bit, same order
    assert np.array_equal(f, gold[name + "_faces"])


@pytest.mark.skipif(not mcubes_ref.available(), reason="oracle/_ref/libmcubes_ref.so not built (needs /root/reference)")
def test_restatement_equals_compiled_reference_random():
    rng = np.random.default_rng(5)
    for shape in [(2, 2, 2), (1, 4, 4), (3, 2, 9), (16, 16, 16), (33, 9, 17)]:
        for kind in range(3):
            vol = rng.standard_normal(shape)
            if kind == 1:
                vol = np.round(vol * 2)                       # ties: f1 == f2, v == iso
            if kind == 2:
                vol = np.pad(vol, 1, "constant", constant_values=-1e6)
            a = mcubes_ref.marching_cubes(vol, 0.0)
            b = co.marching_cubes(vol, 0.0)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (shape, kind)


def test_extract_mesh_is_closed_and_in_box(gold):
    """Generator3D.extract_mesh pads with -1e6 "to make sure that mesh is watertight": every edge is shared by exactly two
    faces, and the vertices land inside the padded unit box."""
    g = np.linspace(-0.55, 0.55, 33)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    occ = (0.45 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)) * 20
    v, f = co.extract_mesh(occ, threshold=0.2, padding=0.1)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    assert np.abs(v).max() <= 0.55 + 1e-12
    r = np.linalg.norm(v, axis=1)
    iso_r = 0.45 - (np.log(0.2) - np.log(0.8)) / 20
    assert np.abs(r - iso_r).max() < 2e-3                      # the iso-surface of the field is a sphere of that radius


def test_case_tables_are_consistent():
    """Generated tables: the edges a case's triangles use are exactly the edges that carry a vertex; complementary
    cases cut the same edges; 820 triangles in total (Lorensen-Cline)."""
    txt = open(os.path.join(ROOT, "if-defense_b200", "csrc", "mc_tables.inc")).read()
    tri = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ull", txt)]
    masks = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{3})\b", txt.split("IFD_MC_EDGE_MASKS")[2])]
    assert len(tri) == 256 and len(masks) == 256
    total = 0
    for c in range(256):
        used = 0
        n = 0
        for s in range(16):
            e = tri[c] >> 4 * s & 0xF
            if e == 0xF:
                assert all((tri[c] >> 4 * t & 0xF) == 0xF for t in range(s, 16))
                break
            assert e < 12
            used |= 1 << e
            n += 1
        assert n % 3 == 0 and used == masks[c]
        assert masks[c] == masks[255 - c]
        total += n // 3
    assert total == 820 and masks[0] == 0


def test_sample_surface_restatement_stays_on_the_mesh():
    v = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    f = np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]])
    u = np.random.default_rng(0).random((4000, 3))
    p, fi = co.sample_surface(v, f, u)
    for t in range(4):
        a, b, c = v[f[t]]
        n = np.cross(b - a, c - a)
        assert np.abs((p[fi == t] - a) @ n).max() < 1e-12
    share = np.bincount(fi, minlength=4) / 4000.0
    area = np.array([0.5, 0.5, 0.5, np.sqrt(3) / 2])
    assert np.abs(share - area / area.sum()).max() < 0.03
    assert (p >= -1e-12).all() and (p.sum(axis=1) <= 1 + 1e-12).all()


MISE_CASES = ["s0", "s1", "s2", "s3", "d0", "big0"]


@pytest.mark.parametrize("name", MISE_CASES)
def test_mise_restatement_equals_reference_fixture(name):
    """tests/golden/mise.npz holds what the reference's own Cython MISE did on these fields: points per round, dense grid."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "mise.npz"))
    r0, depth, kind = (int(x) for x in g[name + "_cfg"])
    res = r0 << depth
    dense, rounds = mise_port.mise_loop(lambda p: mise_port.analytic_field(p, res, kind), r0, depth, float(g[name + "_thr"]))
    assert rounds == list(g[name + "_rounds"])
    assert hashlib.sha256(np.ascontiguousarray(dense).tobytes()).hexdigest() == str(g[name + "_sha256"])
    if name + "_dense" in g:
        assert np.array_equal(dense, g[name + "_dense"].astype(np.float64))


def test_oracle_generate_from_latent_equals_reference_generator3d():
    """tests/golden/generator.npz records a run of the reference's OWN Generator3D.generate_from_latent (its MISE, its
    marching cubes, its decoder, CPU).  Fed with the values that run evaluated, the oracle's MISE asks for exactly the same
    points in every round and hands the same lattice to marching cubes; the oracle's extract_mesh returns the same
    vertices and faces, bit for bit and in order.  The oracle decoder reproduces the logged values to fp32 rounding."""
    import torch
    from ifdefense_b200 import models
    from oracle import torch_port as tp
    g = np.load(os.path.join(GOLDEN, "generator.npz"))
    res, thr = 32, np.log(0.2) - np.log(0.8)
    ijk = np.rint((g["points_f"].astype(np.float64) / 1.1 + 0.5) * res).astype(np.int64)
    assert np.abs((1.1 * (ijk.astype(np.float32) / res - 0.5)) - g["points_f"]).max() == 0       # exact inverse of :121-124
    table = {tuple(p): v for p, v in zip(ijk.tolist(), g["values"].tolist())}
    assert len(table) == len(ijk)                                                                # no point evaluated twice

    def lookup(pts):
        return np.array([table[tuple(p)] for p in pts.tolist()])                                 # KeyError = a point the reference never asked for
    dense, rounds = mise_port.mise_loop(lookup, 8, 2, thr)
    assert rounds == list(g["rounds"])
    assert np.array_equal(dense, g["value_grid"].astype(np.float64))
    v, f = co.extract_mesh(dense, threshold=0.2, padding=0.1)
    assert np.array_equal(v, g["verts"]) and np.array_equal(f, g["faces"])
    sd = models.synthetic_state_dict("onet", 0)
    sd["decoder.fc_out.bias"] = sd["decoder.fc_out.bias"] + float(g["bias_shift"])
    got = tp.onet_decode(sd, torch.from_numpy(g["points_f"])[None], torch.from_numpy(g["c"]))[0].numpy()
    assert np.abs(got - g["values"]).max() < 2e-5
