"""-m gpu: ONet decoder (tcgen05 GEMM chain with folded CBN) and ONet-Opt through the C ABI against the
fixtures generated from the reference (tests/golden/onet.npz)."""
import numpy as np
import pytest
import torch

from ifdefense_b200 import capi, models, onet as onet_mod, synth
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx(conftest_golden_dir):
    return dict(np.load(conftest_golden_dir + "/onet.npz"))


@pytest.fixture(scope="module")
def dec():
    return onet_mod.ONetDecoder(models.synthetic_state_dict("onet", 0))


@pytest.mark.parametrize("tag", ["b2", "cfg0"])
def test_decode_autograd_seam(fx, dec, tag):
    """generator.model.decode(p, z, c).logits and its gradient w.r.t. p (ONet/opt_defense.py:212)."""
    p = dev(fx[tag + "/p0"]).requires_grad_()
    c = dev(fx[tag + "/c"])
    logits = dec.decode(p, torch.empty(p.shape[0], 0), c).logits
    (logits * dev(fx[tag + "/gl"])).sum().backward()
    lerr = np.abs(logits.detach().cpu().numpy() - fx[tag + "/logits"]).max()
    gref = fx[tag + "/grad_p"]
    gerr = np.abs(p.grad.cpu().numpy() - gref)
    print(tag, "logit err", lerr, "grad err/max|g|", gerr.max() / np.abs(gref).max(), "p99.9", np.quantile(gerr, 0.999) / np.abs(gref).max())
    assert lerr < 2e-5
    # fp32-class with a higher ReLU-kink flip rate than fp32 SIMT (3xTF32): tight on almost all, loose on the rest
    assert np.quantile(gerr, 0.99) < 2e-5 * np.abs(gref).max() and gerr.max() < 0.05 * np.abs(gref).max()


def test_onet_opt_short_horizon(fx, dec):
    rest = onet_mod.ONetRestorer(dec, threshold=0.2, lr=1e-3)
    for n_it, key, frac in ((0, "b2/xyz_0", 1.0), (1, "b2/xyz_1", 0.99), (9, "b2/xyz_9", 0.95)):
        out = rest.optimize_points(dev(fx["b2/p0"]), None, dev(fx["b2/c"]), rep_weight=500., iterations=n_it, normalize=False)
        d = np.abs(out - fx[key])
        assert (d < 5e-6).mean() >= frac and np.median(d) < 1e-7, (n_it, (d < 5e-6).mean())
    out = rest.optimize_points(dev(fx["b2/p0"]), None, dev(fx["b2/c"]), rep_weight=500., iterations=19, printing=True)
    d = np.abs(out - fx["b2/final_normalized"])
    assert (d < 1e-4).mean() > 0.95 and np.median(d) < 1e-6
    assert rest.last_stats.shape == (1, 4) and np.isfinite(rest.last_stats).all()
    out2 = rest.optimize_points(dev(fx["b2/p0"]), None, dev(fx["b2/c"]), rep_weight=500., iterations=19)
    assert np.array_equal(out, out2)                                   # bitwise reproducible


def test_chain_launch_equals_layer_by_layer():
    """The ten layers of a direction as one launch of the GEMM engine (tc::gemm_chain_kernel, per-tile hand-over inside a CTA)
    give the bits of ten launches -- the arithmetic per tile is the same, only the launch boundaries go.  Shapes with one row
    tile per CTA (the hand-over has no later tile to hide behind), several, and a ragged last tile."""
    L = capi.lib()
    for B, K in ((1, 96), (2, 1024), (24, 1024), (3, 1000)):
        case = synth.make_onet_case(B, K=K, seed=B)
        rest = onet_mod.ONetRestorer(onet_mod.ONetDecoder(case.sd), threshold=0.2, lr=1e-3)
        outs = []
        for chain in (1, 0):
            L.ifd_test_hook(8, chain)
            try:
                outs.append(rest.optimize_points(case.p0.cuda(), None, case.c.cuda(), rep_weight=500., iterations=5, normalize=False))
            finally:
                L.ifd_test_hook(8, 1)
        assert np.isfinite(outs[0]).all() and np.array_equal(outs[0], outs[1]), (B, K)


def test_config0_on_gpu(fx, dec):
    """BASELINE.json configs[0] (1 cloud x 1024 points, 20 iterations) against the reference's CPU result."""
    rest = onet_mod.ONetRestorer(dec)
    out = rest.optimize_points(dev(fx["cfg0/p0"]), None, dev(fx["cfg0/c"]), rep_weight=500., iterations=20)
    d = np.abs(out - fx["cfg0/final_normalized"])
    assert out.shape == (1, 1024, 3) and (d < 1e-4).mean() > 0.95 and np.median(d) < 1e-6
    np.testing.assert_allclose(np.linalg.norm(out, axis=2).max(), 1.0, rtol=1e-6)


def test_full_size_onet_opt_properties():
    """B=16 x 1024 points, 41 steps: finite, centred, unit max-norm; ragged B*K (not a multiple of 128 rows)."""
    case = synth.make_onet_case(16, K=1024, seed=1, device="cuda")
    d = onet_mod.ONetDecoder(case.sd)
    rest = onet_mod.ONetRestorer(d)
    out = rest.optimize_points(case.p0.cuda(), None, case.c.cuda(), rep_weight=500., iterations=40)
    assert out.shape == (16, 1024, 3) and np.isfinite(out).all() and np.abs(out.mean(1)).max() < 1e-5
    small = rest.optimize_points(case.p0[:3, :1000].cuda().contiguous(), None, case.c[:3].cuda(), rep_weight=500., iterations=3)
    assert small.shape == (3, 1000, 3) and np.isfinite(small).all()


def test_eval_points_and_dense_grid_vs_oracle(dec):
    """The occupancy evaluation of ONet-Mesh (Generator3D.eval_points, generation.py:138-158): chunked forward decode for
    one shape, and the dense (R+1)^3 lattice in the reference's coordinates, against the oracle decoder on the CPU."""
    from oracle import torch_port as tp
    sd = models.synthetic_state_dict("onet", 0)
    case = synth.make_onet_case(1, K=64, seed=2)
    c = case.c[:1]
    p = (torch.rand(2500, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 1.1
    got = dec.eval_points(p, None, c, points_batch_size=1000)                   # 3 chunks, the last one ragged
    want = tp.onet_decode(sd, p[None], c)[0]
    assert got.shape == (2500,) and got.device.type == "cpu"
    assert (got - want).abs().max().item() < 2e-5
    grid = dec.eval_dense_grid(c, resolution=16, points_batch_size=2000)
    assert grid.shape == (17, 17, 17)
    ax = 1.1 * (torch.arange(17, dtype=torch.float32) / 16 - 0.5)
    pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).view(1, -1, 3)
    want = tp.onet_decode(sd, pts, c)[0].view(17, 17, 17)
    assert (grid.cpu() - want).abs().max().item() < 2e-5
