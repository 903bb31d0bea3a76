"""-m gpu: ConvONet decode and the restoration loop through the C ABI against the golden fixtures (generated
from the real reference), the oracle, and size-independent properties at BASELINE.json's full size."""
import ctypes

import numpy as np
import pytest
import torch

from ifdefense_b200 import capi, convonet, synth
from tests.gpu_util import dev, run_opt

pytestmark = pytest.mark.gpu
D, F32 = ctypes.c_double, ctypes.c_float


@pytest.fixture(scope="module")
def dec(conv_sd):
    return convonet.ConvONetDecoder(conv_sd, padding=0.1)


@pytest.fixture(scope="module")
def planes(conv_planes):
    return convonet.planes_to_channels_last({k: v.cuda() for k, v in conv_planes.items()})


def test_layout_conversion_exact(conv, planes):
    want = conv["planes_nchw"].transpose(0, 1, 3, 4, 2)
    assert planes.shape == (3, 2, 64, 64, 32) and np.array_equal(planes.cpu().numpy(), want)


def test_packed_weights_match_fixture(conv, dec):
    assert np.array_equal(dec.blob.cpu().numpy(), conv["dec_blob"])


@pytest.mark.parametrize("pk,lk,gk", [("p0", "logits", "grad_p"), ("clamp_p", "clamp_logits", "clamp_grad_p")])
def test_decode_autograd_seam(conv, dec, conv_planes, pk, lk, gk):
    """generator.model.decode(p, c).logits, differentiable w.r.t. p (opt_defense.py:212)."""
    c = {k: v.cuda() for k, v in conv_planes.items()}
    p = dev(conv[pk]).requires_grad_()
    logits = dec.decode(p, c).logits
    (logits * dev(conv["gl"])).sum().backward()
    assert np.abs(logits.detach().cpu().numpy() - conv[lk]).max() < 2e-6
    assert np.abs(p.grad.cpu().numpy() - conv[gk]).max() < 2e-6 * np.abs(conv[gk]).max()


def test_loop_short_horizon_vs_reference(conv, dec, planes):
    """The north_star tolerance (1e-4 max-abs on xyz) holds with two orders of magnitude to spare over the
    horizon where the reference's own trajectory is reproducible ("Parity tiers")."""
    for n_steps, tol in ((1, 1e-6), (2, 1e-6), (10, 5e-6), (20, 2e-5)):
        x, _ = run_opt(dec, planes, conv["p0"], n_steps, decode_kernel=2)            # fp32 kernels
        assert np.abs(x - conv["trace/xyz_%d" % (n_steps - 1)]).max() < tol, n_steps
    x, _ = run_opt(dec, planes, conv["p0"], 20, normalize=1, decode_kernel=2)
    assert np.abs(x - conv["final_20_normalized"]).max() < 1e-4
    x, _ = run_opt(dec, planes, conv["p0"], 20, normalize=1)                          # production default (tensor cores)
    assert np.abs(x - conv["final_20_normalized"]).max() < 1e-4                       # north_star's tolerance, max-abs
    for n_steps, tol in ((1, 1e-6), (2, 1e-6), (10, 5e-6), (20, 2e-5)):               # and the fp32 kernels' own bounds
        x, _ = run_opt(dec, planes, conv["p0"], n_steps)                              # (measured: 5e-10 / 3e-8 / 6e-8 / 2.9e-7)
        assert np.abs(x - conv["trace/xyz_%d" % (n_steps - 1)]).max() < tol, n_steps


def test_late_state_single_step(conv, dec, planes):
    """Resume from the reference's own (xyz, m, v) after 150 steps and take step 151."""
    m, v = dev(conv["trace/late_m"]).clone(), dev(conv["trace/late_v"]).clone()
    x, _ = run_opt(dec, planes, conv["trace/late_xyz"], 1, m=m, v=v, step0=150, decode_kernel=2)
    assert np.abs(x - conv["trace/late_xyz_next"]).max() < 1e-6


def test_loop_equals_host_instantiation(conv, dec, planes, mathcheck):
    """Device kernels == the same __host__ __device__ arithmetic run serially on the CPU, over 30 steps."""
    pl = np.ascontiguousarray(conv["planes_nchw"].transpose(0, 1, 3, 4, 2))
    xyz = np.ascontiguousarray(conv["p0"]).copy()
    B, K, _ = xyz.shape
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    mathcheck.mc_convonet_opt(P(conv["dec_blob"]), P(pl), P(xyz), B, K, 64, 5, 30, B, 5, D(1e-3), D(0.9), D(0.999), D(1e-8), D(0.2),
                              D(500.), D(0.07), D(0.03), D(1e-12), D(0.1), 0, None, 0, None)
    x, _ = run_opt(dec, planes, conv["p0"], 30, decode_kernel=2)
    assert np.abs(x - xyz).max() < 5e-5


def test_production_kernels_equal_first_generation(conv, dec, planes):
    """The fp32 production-grade loop (decode v2: cooperative gather, 2 points/thread, FFMA2; warm-started kNN)
    against the first-generation kernels (thread-per-point decode, from-scratch kNN every step).  The kNN result is
    identical by construction and the MLP accumulates in the same order, so only the reduction order of the
    gather backward differs: agreement is at rounding level until the trajectory's own sensitivity kicks in."""
    for n_steps, tol in ((1, 2e-7), (5, 1e-6), (25, 2e-5)):
        a, _ = run_opt(dec, planes, conv["p0"], n_steps, decode_kernel=2)
        b, _ = run_opt(dec, planes, conv["p0"], n_steps, decode_kernel=1)
        assert np.abs(a - b).max() < tol, n_steps
    case = synth.make_case(3, K=1024, seed=5, device="cuda")        # K = 1024, B*K not a multiple of the 512-point tile
    d3 = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    a, sa = run_opt(d3, pl, case.p0[:, :1000].contiguous(), 12, decode_kernel=2, stats=True)
    b, sb = run_opt(d3, pl, case.p0[:, :1000].contiguous(), 12, decode_kernel=1, stats=True)
    assert np.abs(a - b).max() < 2e-6
    np.testing.assert_allclose(sa, sb, rtol=1e-6)


def test_tensor_core_selftest():
    """tcgen05 / TMEM / 3xTF32 plumbing: D = A . B^T for one 128 x 32 x 32 tile against float64."""
    rng = np.random.default_rng(0)
    A = (rng.normal(size=(128, 32)) * rng.uniform(0.01, 3.0, size=(128, 1))).astype(np.float32)
    Bm = rng.normal(size=(32, 32)).astype(np.float32) / 5.0
    Ad, Bd = dev(A), dev(Bm)
    D = torch.zeros((128, 32), dtype=torch.float32, device="cuda")
    capi.check(capi.lib().ifd_selftest_umma(capi.ptr(Ad), capi.ptr(Bd), capi.ptr(D), capi.stream()))
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ Bm.astype(np.float64).T
    scale = (np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64).T)
    err = np.abs(D.cpu().numpy() - want) / scale
    assert err.max() < 2e-6, err.max()         # fp32-class: a plain fp32 FMA chain gives ~1e-7..1e-6 here


def bce_grad(dec, planes, xyz, kernel, B_ref=None):
    x = dev(xyz)
    B, K, _ = x.shape
    L = capi.lib()
    ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(B, K), dtype=torch.uint8, device="cuda")
    g = torch.empty_like(x)
    capi.check(L.ifd_convonet_decode_bce_grad(capi.ptr(planes), capi.ptr(dec.blob), capi.ptr(x), B, K, planes.shape[2], 32, 32, dec.dims[2],
                                              0.1, 0.2, B if B_ref is None else B_ref, kernel, capi.ptr(g), capi.ptr(ws),
                                              ws.numel(), capi.stream()))
    torch.cuda.synchronize()
    return g.cpu().numpy()


def test_bce_grad_seam_all_kernels(conv, conv_sd, conv_planes, dec, planes):
    """d(K * mean BCE(logit, 0.2))/dxyz from the three decode kernels against torch autograd on the oracle."""
    import torch.nn.functional as F
    from oracle import torch_port as tp
    p = torch.from_numpy(conv["p0"]).requires_grad_()
    lg = tp.convonet_decode(conv_sd, p, conv_planes)
    (F.binary_cross_entropy_with_logits(lg, torch.full_like(lg, 0.2), reduction="none").mean() * p.shape[1]).backward()
    want = p.grad.numpy()
    scale = np.abs(want).max()
    errs = {}
    grads = {}
    for kernel in (1, 2, 5):
        g = grads[kernel] = bce_grad(dec, planes, conv["p0"], kernel)
        errs[kernel] = np.abs(g - want).max() / scale
    print("bce-grad max error / max|grad| per kernel:", errs)
    assert errs[1] < 2e-6 and errs[2] < 2e-6
    assert errs[5] < 1e-5                         # 3xTF32 tensor-core path: fp32-class (vs a float64 evaluation: 2.8e-6, the fp32 kernels 3.0e-6)


@pytest.mark.parametrize("nb", [1, 2, 3, 8])
def test_decoder_depths_other_than_five(conv, conv_planes, planes, nb):
    """LocalDecoder with n_blocks != 5 (the stage table of decode v5 -- which layers ride with which round -- depends on it):
    BCE gradient of the tensor-core kernel and of the fp32 kernel against torch autograd on the oracle, and a short loop."""
    import torch.nn.functional as F
    from oracle import torch_port as tp
    g = torch.Generator().manual_seed(100 + nb)
    rnd = lambda *shape: torch.randn(shape, generator=g) / (shape[-1] ** 0.5)
    sd = {"decoder.fc_p.weight": rnd(32, 3), "decoder.fc_p.bias": rnd(32) * 0.3, "decoder.fc_out.weight": rnd(1, 32),
          "decoder.fc_out.bias": rnd(1) * 0.1}
    for i in range(nb):
        sd["decoder.fc_c.%d.weight" % i], sd["decoder.fc_c.%d.bias" % i] = rnd(32, 32), rnd(32) * 0.3
        for fc in ("fc_0", "fc_1"):
            sd["decoder.blocks.%d.%s.weight" % (i, fc)], sd["decoder.blocks.%d.%s.bias" % (i, fc)] = rnd(32, 32), rnd(32) * 0.3
    d = convonet.ConvONetDecoder(sd, padding=0.1)
    assert d.dims == (32, 32, nb)
    p = torch.from_numpy(conv["p0"]).requires_grad_()
    lg = tp.convonet_decode(sd, p, conv_planes, n_blocks=nb)
    (F.binary_cross_entropy_with_logits(lg, torch.full_like(lg, 0.2), reduction="none").mean() * p.shape[1]).backward()
    want = p.grad.numpy()
    scale = np.abs(want).max()
    assert np.abs(bce_grad(d, planes, conv["p0"], 2) - want).max() / scale < 2e-6
    assert np.abs(bce_grad(d, planes, conv["p0"], 5) - want).max() / scale < 1e-5
    a, _ = run_opt(d, planes, conv["p0"], 5, decode_kernel=5)
    b, _ = run_opt(d, planes, conv["p0"], 5, decode_kernel=2)
    assert np.isfinite(a).all() and (np.abs(a - b) < 1e-6).mean() > 0.98


@pytest.mark.parametrize("tck", [5])
def test_tensor_core_decode_kernel(conv, dec, planes, tck):
    """decode v5 (ResNet-MLP on tcgen05, 3xTF32).  Its gradient is as close to a float64 evaluation as the
    fp32 reference's own (test_bce_grad_seam_all_kernels, tools/grad_probe.py), but pre-activations carry
    ~2^-21 instead of ~2^-24 relative noise, so a ReLU sitting within that noise of zero flips its sign mask
    about 8x more often than between two fp32 implementations (about one point in 500 per evaluation on these
    weights).  A flip changes that point's gradient by O(1 %) -- both values are exact gradients of the same
    piecewise-linear function on either side of a kink.  Hence: tight bounds on almost all coordinates, a loose
    bound on the few that crossed a kink."""
    for n_steps, tol, frac in ((1, 1e-6, 1.0), (2, 1e-6, 0.99), (10, 5e-6, 0.97), (20, 2e-5, 0.95)):
        x, _ = run_opt(dec, planes, conv["p0"], n_steps, decode_kernel=tck)
        d = np.abs(x - conv["trace/xyz_%d" % (n_steps - 1)])
        assert (d < tol).mean() >= frac, (n_steps, (d < tol).mean())
        assert d.max() < 2.5e-3 * n_steps ** 0.5 and np.median(d) < 1e-7
    a, sa = run_opt(dec, planes, conv["p0"], 5, decode_kernel=tck, stats=True)
    b, sb = run_opt(dec, planes, conv["p0"], 5, decode_kernel=2, stats=True)
    assert (np.abs(a - b) < 1e-6).mean() > 0.98
    np.testing.assert_allclose(sa, sb, rtol=1e-5)
    m, v = dev(conv["trace/late_m"]).clone(), dev(conv["trace/late_v"]).clone()
    x, _ = run_opt(dec, planes, conv["trace/late_xyz"], 1, m=m, v=v, step0=150, decode_kernel=tck)
    d = np.abs(x - conv["trace/late_xyz_next"])
    assert (d < 1e-6).mean() > 0.99 and d.max() < 1e-4
    case = synth.make_case(3, K=1024, seed=5, device="cuda")         # ragged last tile
    d3 = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    a, _ = run_opt(d3, pl, case.p0[:, :1000].contiguous(), 12, decode_kernel=tck)
    b, _ = run_opt(d3, pl, case.p0[:, :1000].contiguous(), 12, decode_kernel=2)
    assert (np.abs(a - b) < 5e-6).mean() > 0.97 and np.isfinite(a).all()
    a2, _ = run_opt(d3, pl, case.p0[:, :1000].contiguous(), 12, decode_kernel=tck)
    assert np.array_equal(a, a2)                                     # bitwise reproducible


def test_tensor_core_201_steps_statistical(conv, dec, planes):
    x, st = run_opt(dec, planes, conv["p0"], 201, stats=True, decode_kernel=5)
    ref = conv["final_201_raw"]
    d = np.abs(x - ref)
    # measured on a B200 (profiles/r02_parity_record.json, fixture 2 x 256): median 5.5e-5, p99 1.6e-3, max 1.9e-2, 62.5 % within
    # 1e-4 -- the same as the fp32 kernels (4.5e-5 / 1.2e-3 / 5.5e-3 / 63 %): the 201-step trajectory amplifies ANY rounding difference
    assert np.isfinite(x).all() and np.median(d) < 1.5e-4 and np.quantile(d, 0.99) < 5e-3 and d.max() < 0.06 and (d <= 1e-4).mean() > 0.5
    np.testing.assert_allclose(st[0], conv["stats_201"][0], rtol=2e-5)
    np.testing.assert_allclose(st[1:], conv["stats_201"][1:], rtol=0.05)


def test_201_steps_statistical_parity(conv, dec, planes):
    """At 201 steps the reference does not reproduce itself (8-thread vs 1-thread CPU runs differ by 3e-2
    max-abs, median 1.5e-4): parity is distributional, measured against that noise floor."""
    x, st = run_opt(dec, planes, conv["p0"], 201, stats=True)
    ref = conv["final_201_raw"]
    d = np.abs(x - ref)
    assert np.isfinite(x).all()
    # thresholds = the recorded distribution (profiles/r02_parity_record.json: median 5.5e-5, p99 1.6e-3, max 1.9e-2) + margin
    assert np.median(d) < 1.5e-4 and np.quantile(d, 0.99) < 5e-3 and d.max() < 0.06 and (d <= 1e-4).mean() > 0.5
    # the printed diagnostics: exact at iteration 0, statistically equal later
    np.testing.assert_allclose(st[0], conv["stats_201"][0], rtol=2e-5)
    np.testing.assert_allclose(st[1:], conv["stats_201"][1:], rtol=0.05)
    # same point set up to noise: symmetric chamfer distance far below the point spacing (~0.03)
    a, b = torch.from_numpy(x).cuda(), torch.from_numpy(ref).cuda()
    cd = torch.cdist(a, b)
    assert float(cd.min(2)[0].mean()) < 5e-3 and float(cd.min(1)[0].mean()) < 5e-3


def test_bitwise_determinism_and_batch_split(dec):
    """Run-to-run bitwise reproducibility, and independence from how a reference batch is split across calls
    (the multi-GPU sharding contract): clouds 0-1 and 2-3 restored separately with B_ref = 4 == all four."""
    case = synth.make_case(4, K=512, seed=2, device="cuda")
    d4 = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    a, _ = run_opt(d4, pl, case.p0, 60, B_ref=4, normalize=1)
    b, _ = run_opt(d4, pl, case.p0, 60, B_ref=4, normalize=1)
    assert np.array_equal(a, b)
    lo, _ = run_opt(d4, pl[:, :2].contiguous(), case.p0[:2], 60, B_ref=4, normalize=1)
    hi, _ = run_opt(d4, pl[:, 2:].contiguous(), case.p0[2:], 60, B_ref=4, normalize=1)
    assert np.array_equal(np.concatenate([lo, hi]), a)
    c, _ = run_opt(d4, pl, case.p0, 60, B_ref=2, normalize=1)          # and B_ref does matter
    assert not np.array_equal(a, c)


def test_reference_signature_and_host_seam(conv, dec, conv_planes):
    """optimize_points(opt_points, z, c, rep_weight, iterations, printing) -> numpy [B,K,3]; the host-buffer
    entry point gives the same bits."""
    rest = convonet.Restorer(dec, threshold=0.2, lr=1e-3, decode_kernel=2)
    c = {k: v.cuda() for k, v in conv_planes.items()}
    out = rest.optimize_points(dev(conv["p0"]), None, c, rep_weight=500., iterations=19, printing=True)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == conv["p0"].shape
    assert np.abs(out - conv["final_20_normalized"]).max() < 1e-4
    assert rest.last_stats.shape == (1, 4)
    host = rest.optimize_points_host(conv["p0"], conv["planes_nchw"], rep_weight=500., iterations=19)
    assert np.array_equal(host, out)
    many = rest.optimize_points_host_many([conv["p0"]] * 3, [conv["planes_nchw"]] * 3, rep_weight=500., iterations=19)
    assert len(many) == 3 and all(np.array_equal(m, host) for m in many)      # pipelined batch loop: same bits
    pinned = [torch.empty(conv["p0"].shape, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    into = rest.optimize_points_host_many([conv["p0"]] * 2, [conv["planes_nchw"]] * 2, rep_weight=500., iterations=19, out=pinned)
    assert all(a is b for a, b in zip(into, pinned)) and all(np.array_equal(m, host) for m in pinned)
    with pytest.raises(RuntimeError, match="one output array per batch"):
        rest.optimize_points_host_many([conv["p0"]] * 2, [conv["planes_nchw"]] * 2, out=pinned[:1])
    out0 = rest.optimize_points(dev(conv["p0"]), None, c, rep_weight=0., iterations=19)     # rep_weight == 0 branch
    assert np.isfinite(out0).all() and not np.array_equal(out0, out)


def test_device_batches_equal_single_calls(conv, dec, planes):
    """ifd_convonet_opt_batches (loops side by side on internal streams, one-CTA tail) gives the bits of one call per batch
    (a loop alone: cluster-pair tail)."""
    L = capi.lib()
    want, _ = run_opt(dec, planes, conv["p0"], 20, normalize=1)
    xs = [dev(conv["p0"]).clone() for _ in range(3)]
    B, K, _ = xs[0].shape
    one = (L.ifd_convonet_opt_workspace_bytes(B, K) + 255) // 256 * 256
    ws = torch.empty(L.ifd_convonet_opt_batches_workspace_bytes(B, K), dtype=torch.uint8, device="cuda")
    assert ws.numel() % one == 0 and ws.numel() >= 2 * one
    P = capi.default_params(n_steps=20, B_ref=B, normalize_out=1)
    pp = (ctypes.c_void_p * 3)(*[planes.data_ptr()] * 3)
    xp = (ctypes.c_void_p * 3)(*[x.data_ptr() for x in xs])
    capi.check(L.ifd_convonet_opt_batches(3, pp, capi.ptr(dec.blob), xp, B, K, 64, 32, 32, 5, ctypes.byref(P), capi.ptr(ws),
                                          ws.numel(), capi.stream()), "ifd_convonet_opt_batches")
    torch.cuda.synchronize()
    for x in xs:
        assert np.array_equal(x.cpu().numpy(), want)
    with pytest.raises(RuntimeError, match="workspace must hold"):
        capi.check(L.ifd_convonet_opt_batches(3, pp, capi.ptr(dec.blob), xp, B, K, 64, 32, 32, 5, ctypes.byref(P), capi.ptr(ws),
                                              one, capi.stream()))


def test_full_size_properties():
    """BASELINE.json config 2: B=64 x 1024 points, 201 steps.  Size-independent invariants of the output:
    centred, unit max-norm, finite, loss decreased, surface constraint approached."""
    case = synth.make_case(64, K=1024, seed=0, device="cuda")
    d = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    capi.lib().ifd_launch_count(1)
    x, st = run_opt(d, pl, case.p0, 201, normalize=1, stats=True)
    launches = capi.lib().ifd_launch_count(0)
    assert launches >= 201 * 2        # decode + fused per-cloud tail per Adam step
    assert x.shape == (64, 1024, 3) and np.isfinite(x).all()
    assert np.abs(x.mean(1)).max() < 1e-5
    np.testing.assert_allclose(np.linalg.norm(x, axis=2).max(1), 1.0, rtol=1e-6)
    assert st[2, 0] < st[0, 0] and st[2, 2] < st[0, 2]                 # total and repulsion loss went down
    raw, _ = run_opt(d, pl, case.p0, 201, normalize=0)
    assert np.abs(raw - case.p0.numpy()).max() < 0.25                  # |step| <= ~lr per iteration


def test_reference_default_batch_192_and_ragged_last_batch():
    """The reference's default --batch_size=192 (opt_defense.py:40) and its ragged last batch of 164 (2468 = 12 x 192 +
    164): more CTAs than SMs for both kernels; restoring the batch in one call == in three slices with B_ref."""
    for Bfull in (192, 164):
        case = synth.make_case(Bfull, K=1024, seed=3, device="cuda")
        d = convonet.ConvONetDecoder(case.sd)
        pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
        full, _ = run_opt(d, pl, case.p0, 8, normalize=1)
        assert full.shape == (Bfull, 1024, 3) and np.isfinite(full).all()
        cuts = [0, Bfull // 3, 2 * Bfull // 3 + 1, Bfull]
        parts = [run_opt(d, pl[:, a:b].contiguous(), case.p0[a:b], 8, B_ref=Bfull, normalize=1)[0] for a, b in zip(cuts, cuts[1:])]
        assert np.array_equal(np.concatenate(parts), full)


def test_large_clouds_take_the_first_generation_tail_and_edge_step_counts(dec, planes, conv):
    """K > 1024 points per cloud: the fused tail does not apply, the loop switches to the brute-force kNN + separate Adam
    kernels by itself (same result as asking for them); n_steps = 0 leaves the points untouched."""
    case = synth.make_case(2, K=1536, seed=4, device="cuda")
    d = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    a, _ = run_opt(d, pl, case.p0, 6)
    b, _ = run_opt(d, pl, case.p0, 6, tail_kernel=1)
    assert a.shape == (2, 1536, 3) and np.isfinite(a).all() and np.array_equal(a, b)
    x0, _ = run_opt(dec, planes, conv["p0"], 0)
    assert np.array_equal(x0, conv["p0"])


def test_errors(dec, planes, conv):
    x = dev(conv["p0"])
    L = capi.lib()
    P = capi.default_params()
    with pytest.raises(RuntimeError, match="workspace"):
        capi.check(L.ifd_convonet_opt(capi.ptr(planes), capi.ptr(dec.blob), capi.ptr(x), None, None, 2, 256, 64, 32, 32, 5,
                                      ctypes.byref(P), None, capi.ptr(x), 16, capi.stream()))
    with pytest.raises(RuntimeError, match="c_dim = hidden_size = 32"):
        capi.check(L.ifd_convonet_decode_fwd(capi.ptr(planes), capi.ptr(dec.blob), capi.ptr(x), 2, 256, 64, 64, 64, 5, 0.1,
                                             capi.ptr(x), capi.stream()))


def test_large_batch_runs_as_side_by_side_parts_with_the_same_bits(conv, dec):
    """Restorer cuts a batch of >= 96 clouds into equal parts of at most 64 whose loops run next to each other (B_ref = the
    batch; one-CTA tail in that mode): same bits as the single loop (cluster-pair tail)."""
    reps = 128 // conv["p0"].shape[0] + 1
    p0 = torch.from_numpy(np.concatenate([conv["p0"]] * reps)[:128].copy())
    p0 += torch.randn(p0.shape, generator=torch.Generator().manual_seed(1)) * 1e-3
    c = {k: dev(np.concatenate([conv["planes_nchw"][i]] * reps)[:128]) for i, k in enumerate(("xz", "xy", "yz"))}
    one = convonet.Restorer(dec, side_by_side=False).optimize_points(p0, None, c, rep_weight=500., iterations=8)
    r = convonet.Restorer(dec)
    assert (r._parts(128), r._parts(192), r._parts(164), r._parts(64)) == (2, 3, 4, 1)
    for n in (True, 4, 2):
        got = convonet.Restorer(dec, side_by_side=n).optimize_points(p0, None, c, rep_weight=500., iterations=8)
        assert np.array_equal(got, one), n


def test_graph_replay_equals_direct_launches(conv, dec, planes):
    """The production loop replays one cached CUDA graph whose kernels read their buffers from a device-side job record
    (ifd_convonet_opt: fresh run, default kernels, no diagnostics).  Same bits as the ~400 direct launches; the graph is
    reused for other buffers, other streams get their own record."""
    L = capi.lib()
    L.ifd_launch_count(1)
    a, _ = run_opt(dec, planes, conv["p0"], 20, normalize=1)
    n_graph = L.ifd_launch_count(1)
    L.ifd_test_hook(3, 0)
    try:
        b, _ = run_opt(dec, planes, conv["p0"], 20, normalize=1)
        n_direct = L.ifd_launch_count(1)
    finally:
        L.ifd_test_hook(3, 1)
    assert np.array_equal(a, b)
    assert n_graph == n_direct + 1                       # the same kernels + the job-record write
    assert np.abs(a - conv["final_20_normalized"]).max() < 1e-4
    # replays on other buffers (new xyz / workspace allocations, a copy of the planes) and on another stream
    pl2 = planes.clone()
    c, _ = run_opt(dec, pl2, conv["p0"].copy(), 20, normalize=1)
    assert np.array_equal(a, c)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        d, _ = run_opt(dec, pl2, conv["p0"], 20, normalize=1)
    assert np.array_equal(a, d)
    # two replays back to back on one stream, every buffer of both alive at the same time (distinct addresses: a kernel that
    # kept a capture-time pointer instead of the record's would restore the wrong cloud -- nvcc 12.9 once compiled
    # cloud_step_kernel that way, and single replays on recycled allocations could not see it)
    C, H, nb = dec.dims
    P = capi.default_params(n_steps=20, B_ref=2, normalize_out=1)
    with torch.cuda.stream(side):
        xa, xb = dev(conv["p0"]).clone(), dev(conv["p0"] * 0.97).clone()
        wsb = L.ifd_convonet_opt_workspace_bytes(2, 256)
        wsa, wsb_t = torch.empty(wsb, dtype=torch.uint8, device="cuda"), torch.empty(wsb, dtype=torch.uint8, device="cuda")
        for x, ws in ((xa, wsa), (xb, wsb_t)):
            capi.check(L.ifd_convonet_opt(capi.ptr(pl2), capi.ptr(dec.blob), capi.ptr(x), None, None, 2, 256, 64, C, H, nb,
                                          ctypes.byref(P), None, capi.ptr(ws), wsb, side.cuda_stream))
        side.synchronize()
    assert np.array_equal(xa.cpu().numpy(), a)
    wb, _ = run_opt(dec, planes, conv["p0"] * 0.97, 20, normalize=1)
    assert np.array_equal(xb.cpu().numpy(), wb)
    e, _ = run_opt(dec, planes, conv["p0"], 7)           # another step count: another graph
    L.ifd_test_hook(3, 0)
    try:
        f, _ = run_opt(dec, planes, conv["p0"], 7)
    finally:
        L.ifd_test_hook(3, 1)
    assert np.array_equal(e, f)

