"""-m gpu: the ONet-Mesh tail through the C ABI -- marching cubes (bit for bit and in the reference's order against the
reference-generated fixture and the oracle restatement), extract_mesh's padding / box transform, surface sampling, and
the Generator3D / resample_points mirrors."""
import ctypes
import os
import time

import numpy as np
import pytest
import torch

from ifdefense_b200 import capi, mesh, models, onet as onet_mod, synth
from oracle import c_oracle as co
from oracle import mise_port

from .conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "mesh.npz"))


@pytest.mark.parametrize("name", ["normal", "ints", "ellipsoid_padded"])
def test_marching_cubes_equals_reference_fixture(gold, name):
    v, f = mesh.marching_cubes(gold[name + "_vol"], float(gold[name + "_iso"]))
    assert v.dtype == np.float64 and f.dtype == np.int64
    assert np.array_equal(v, gold[name + "_verts"]) and np.array_equal(f, gold[name + "_faces"])


def test_marching_cubes_random_volumes_vs_oracle():
    rng = np.random.default_rng(11)
    for shape in [(2, 2, 2), (1, 5, 5), (2, 3, 40), (31, 33, 35), (64, 20, 7), (40, 40, 40)]:
        for kind in range(4):
            vol = rng.standard_normal(shape)
            if kind == 1:
                vol = np.round(vol * 2)                                        # f1 == f2 midpoints, values == iso
            if kind == 2:
                vol = vol.astype(np.float32)                                   # float32 in: widened exactly
            if kind == 3:
                vol[rng.random(shape) < 0.1] = np.inf
            want = co.marching_cubes(vol.astype(np.float64), 0.25 if kind != 1 else 0.0)
            got = mesh.marching_cubes(vol, 0.25 if kind != 1 else 0.0)
            assert got[0].shape == want[0].shape and got[1].shape == want[1].shape, (shape, kind)
            assert np.array_equal(got[0], want[0], equal_nan=True) and np.array_equal(got[1], want[1]), (shape, kind)


def test_marching_cubes_empty_and_errors():
    v, f = mesh.marching_cubes(np.ones((6, 6, 6)), 0.0)                        # nothing crosses the isovalue
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = mesh.marching_cubes(np.ones((1, 1, 1)), 0.0)                        # no cells at all
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(RuntimeError, match="three-dimensional"):
        mesh.marching_cubes(np.ones((4, 4)), 0.0)
    with pytest.raises(IndexError):
        mesh.sample_surface_device(torch.zeros((0, 3), dtype=torch.float64, device="cuda"),
                                   torch.zeros((0, 3), dtype=torch.int64, device="cuda"), 16)
    L = capi.lib()
    d = torch.zeros(8, dtype=torch.float64, device="cuda")
    n1, n2 = ctypes.c_longlong(), ctypes.c_longlong()
    ws = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    assert L.ifd_mc_count(capi.ptr(d), 1, 2, 2, 2, 0, 0.0, 0.0, capi.ptr(ws), 16, ctypes.byref(n1), ctypes.byref(n2), capi.stream()) != 0
    assert b"workspace too small" in L.ifd_last_error()
    assert L.ifd_mc_count(capi.ptr(d), 7, 2, 2, 2, 0, 0.0, 0.0, capi.ptr(ws), ws.numel(), ctypes.byref(n1), ctypes.byref(n2), capi.stream()) != 0


@pytest.mark.parametrize("n", [33, 129])
def test_extract_mesh_pad_and_box_transform(n):
    """generation.py:165-186 on an analytic field at the shipped lattice size (129^3): the fused padding + transform give
    the bits of np.pad + marching_cubes + the four numpy statements."""
    g = torch.linspace(-0.55, 0.55, n, dtype=torch.float64)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    occ = ((0.4 - torch.sqrt(X ** 2 + 0.7 * Y ** 2 + 1.2 * Z ** 2)) * 15 + 0.3 * torch.sin(9 * X) * torch.cos(7 * Y)).float()
    want_v, want_f = co.extract_mesh(occ.numpy().astype(np.float64), threshold=0.2, padding=0.1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, f = mesh.extract_mesh(occ.cuda(), threshold=0.2, padding=0.1)
    torch.cuda.synchronize()
    print("extract_mesh %d^3: %d verts, %d faces, %.2f ms on the device" % (n, v.shape[0], f.shape[0], 1e3 * (time.perf_counter() - t0)))
    assert np.array_equal(v.cpu().numpy(), want_v) and np.array_equal(f.cpu().numpy(), want_f)


def test_sample_surface_vs_restatement():
    g = np.linspace(-0.55, 0.55, 49)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    occ = (0.42 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)) * 15
    v, f = mesh.extract_mesh(occ, 0.2, 0.1)
    u = np.random.default_rng(3).random((4096, 3))
    u[0] = [0.0, 0.0, 0.0]
    u[1] = [1.0 - 2 ** -53, 1.0 - 2 ** -53, 1.0 - 2 ** -53]
    pts, fi = mesh.sample_surface_device(v, f, 4096, uniforms=u, return_index=True)
    want_p, want_f = co.sample_surface(v.cpu().numpy(), f.cpu().numpy(), u)
    pts, fi = pts.cpu().numpy(), fi.cpu().numpy()
    # the area prefix is np.cumsum's own chain of rounded fp64 additions (area_cumsum_serial_kernel), so every pick lands
    # on the face the restated trimesh algorithm picks and the points are equal bit for bit
    assert np.array_equal(fi, want_f)
    assert np.array_equal(pts, want_p)
    # every sample lies in its triangle
    tri = v.cpu().numpy()[f.cpu().numpy()[fi]]
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    nrm = np.cross(b - a, c - a)
    assert np.abs(((pts - a) * nrm).sum(1)).max() < 1e-12
    r = np.linalg.norm(pts, axis=1)
    assert r.min() > 0.3 and r.max() < 0.6
    # seeded draws through the rng path are reproducible
    p1 = mesh.sample_surface_device(v, f, 1024, rng=np.random.default_rng(9)).cpu().numpy()
    p2 = mesh.sample_surface_device(v, f, 1024, rng=np.random.default_rng(9)).cpu().numpy()
    assert np.array_equal(p1, p2)


def test_generator3d_and_resample_points():
    """reconstruct_mesh / resample_points (ONet/remesh_defense.py:126-171) with the synthetic ONet: the mesh of the decoder's
    own lattice equals the oracle's extraction of that lattice; 1024 points come back on it."""
    sd = models.synthetic_state_dict("onet", 0)
    dec = onet_mod.ONetDecoder(sd)
    case = synth.make_onet_case(1, K=64, seed=2)
    gen = mesh.Generator3D(dec, threshold=0.2, resolution0=16, upsampling_steps=1, padding=0.1)
    grid = gen.value_grid(case.c[:1])
    assert grid.shape == (33, 33, 33)
    v, f = gen.generate_from_latent(None, case.c[:1])
    want_v, want_f = co.extract_mesh(grid.cpu().numpy().astype(np.float64), 0.2, 0.1)
    assert np.array_equal(v.cpu().numpy(), want_v) and np.array_equal(f.cpu().numpy(), want_f)
    gen0 = mesh.Generator3D(dec, threshold=0.2, resolution0=20, upsampling_steps=0)
    assert gen0.value_grid(case.c[:1]).shape == (20, 20, 20)

    model = models.build_onet()
    model.load_state_dict(sd)
    model = model.cuda().eval()
    pc = np.load(os.path.join(GOLDEN, "airplane.npy"))

    def encode(x):
        with torch.no_grad():
            return model.encode_inputs(x)
    out = mesh.resample_points(gen, encode, pc, num_points=1024, input_npoint=300, rng=np.random.default_rng(0))
    assert out.shape == (1024, 3) and np.isfinite(out).all()
    # a generator whose field never crosses the threshold: the reference's fallback (random subset of the input)
    class Flat(mesh.Generator3D):
        def value_grid(self, c):
            return torch.full((9, 9, 9), -5.0, device="cuda")
    out = mesh.resample_points(Flat(dec), encode, pc[:700], num_points=1024, rng=np.random.default_rng(0))
    assert out.shape == (1024, 3) and np.array_equal(out[:700], pc[:700].astype(np.float32)) and (out[700:] == 0).all()


def _run_mise(r0, depth, thr, fn):
    ex = mesh.MISE(r0, depth, thr)
    rounds = []
    pts = ex.query()
    while pts.shape[0] != 0:
        rounds.append(int(pts.shape[0]))
        ex.update(pts, torch.from_numpy(fn(pts.cpu().numpy())).cuda())
        pts = ex.query()
    return ex.to_dense().cpu().numpy(), rounds


@pytest.mark.parametrize("name", ["s0", "s1", "s2", "s3", "d0", "big0", "big1"])
def test_mise_equals_reference_fixture(name):
    """The device MISE against what the reference's own Cython MISE did (tests/golden/mise.npz): the number of points it
    asks for in every round and the dense lattice it hands to marching cubes."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "mise.npz"))
    r0, depth, kind = (int(x) for x in g[name + "_cfg"])
    res = r0 << depth
    dense, rounds = _run_mise(r0, depth, float(g[name + "_thr"]), lambda p: mise_port.analytic_field(p, res, kind))
    assert rounds == list(g[name + "_rounds"])
    assert hashlib.sha256(np.ascontiguousarray(dense).tobytes()).hexdigest() == str(g[name + "_sha256"])


def test_mise_random_fields_vs_oracle_and_errors():
    rng = np.random.default_rng(4)
    for r0, depth in [(3, 1), (5, 2), (4, 3), (7, 0)]:
        n = (r0 << depth) + 1
        # a smooth random field (a few Fourier modes) plus plateaus at the threshold
        ax = np.arange(n) / (n - 1.0)
        X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
        F = sum(rng.standard_normal() * np.cos(2 * np.pi * (rng.integers(0, 3) * X + rng.integers(0, 3) * Y + rng.integers(0, 3) * Z)
                                               + rng.random()) for _ in range(5))
        F[np.abs(F) < 0.1] = 0.0
        fn = lambda p: F[p[:, 0], p[:, 1], p[:, 2]]
        want, want_rounds = mise_port.mise_loop(fn, r0, depth, 0.0)
        got, rounds = _run_mise(r0, depth, 0.0, fn)
        assert rounds == want_rounds and np.array_equal(got, want), (r0, depth)
    ex = mesh.MISE(4, 2, 0.0)
    with pytest.raises(ValueError, match="Point not in grid"):
        ex.update(torch.tensor([[1, 0, 0]]), torch.tensor([0.5]))                 # not a coarse lattice point
    with pytest.raises(ValueError, match="Point not in grid"):
        ex.update(torch.tensor([[0, 0, 99]]), torch.tensor([0.5]))
    assert ex.query().shape == (125, 3)
    q = ex.query()
    assert q.dtype == torch.int64 and len({tuple(r) for r in q.cpu().numpy().tolist()}) == 125 and (q % 4 == 0).all()


def test_generator3d_mise_path_equals_mise_of_the_dense_lattice():
    """generate_from_latent as the reference runs it (MISE refinement, generation.py:113-130) with the ONet decoder: the
    lattice equals the oracle MISE fed from the densely evaluated lattice (a point's logit does not depend on which other
    points are evaluated with it), and far fewer points are evaluated."""
    case = synth.make_onet_case(1, K=64, seed=2)
    c = case.c[:1]
    sd = synth.onet_with_surface(models.synthetic_state_dict("onet", 0), c, occupied=0.3)
    dec = onet_mod.ONetDecoder(sd)
    full = mesh.Generator3D(dec, 0.2, resolution0=8, upsampling_steps=2, dense=True).value_grid(c).cpu().numpy().astype(np.float64)
    gen = mesh.Generator3D(dec, 0.2, resolution0=8, upsampling_steps=2, points_batch_size=700)
    grid = gen.value_grid(c).cpu().numpy()
    thr = np.log(0.2) - np.log(0.8)
    want, rounds = mise_port.mise_loop(lambda p: full[p[:, 0], p[:, 1], p[:, 2]], 8, 2, thr)
    print("MISE rounds", rounds, "of", 33 ** 3)
    assert len(rounds) >= 3 and gen.points_evaluated == sum(rounds) and gen.points_evaluated < 33 ** 3
    assert np.array_equal(grid, want)
    v, f = gen.generate_from_latent(None, c)
    wv, wf = co.extract_mesh(want, 0.2, 0.1)
    assert np.array_equal(v.cpu().numpy(), wv) and np.array_equal(f.cpu().numpy(), wf)


def test_cumsum_is_numpy_cumsum_bit_for_bit():
    """ifd_cumsum_f64 (the area prefix of ifd_sample_surface) == np.cumsum on float64, every bit: random magnitudes over many
    binades, zeros, subnormals, a tie (an addend of exactly half an ulp of the running sum), addends larger than the sum,
    the real use (triangle areas of a marching-cubes mesh), ragged lengths around the block size."""
    L = capi.lib()
    rng = np.random.default_rng(0)
    cases = []
    for n in (1, 2, 31, 1023, 1024, 1025, 5000, 116416):
        cases.append(rng.random(n) * 10.0 ** rng.integers(-12, 3, size=n))
    x = rng.random(4000)
    x[::7] = 0.0
    x[5] = 5e-324
    x[11] = 2.0 ** -1040
    cases.append(x)
    t = np.array([1.0, 2.0 ** -53, 2.0 ** -53, 1.0, 2.0 ** -52, 3 * 2.0 ** -53, 1e30, 1.0, 2.0 ** 60], dtype=np.float64)      # ties, jumps
    cases.append(t)
    cases.append(np.concatenate([np.full(3000, 2.0 ** -30), [1.0], np.full(3000, 2.0 ** -54 * 1.5), np.full(10, 2.0 ** -53)]))
    cases.append(np.sort(rng.random(20000) * 1e-9)[::-1].copy() + 1e-3)
    for a in cases:
        a = np.ascontiguousarray(a, dtype=np.float64)
        want = np.cumsum(a)
        d = torch.from_numpy(a).cuda()
        tot = torch.zeros(1, dtype=torch.float64, device="cuda")
        capi.check(L.ifd_cumsum_f64(capi.ptr(d), a.shape[0], capi.ptr(tot), capi.stream()), "ifd_cumsum_f64")
        got = d.cpu().numpy()
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (a.shape, np.flatnonzero(got != want)[:5])
        assert float(tot.item()) == want[-1]
