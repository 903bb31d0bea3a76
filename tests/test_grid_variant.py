"""The 'grid' (feature volume, trilinear) variant of ConvONet -- what BASELINE.json's north_star describes, although no shipped
config selects it.  Fixture: tests/golden/grid.npz, generated from the reference's own LocalDecoder /
LocalPoolPointnet by tests/golden/make_grid_golden.py.  CPU: oracle and the product's host-instantiated arithmetic against the
fixture; -m gpu: the kernels through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import torch_port as tp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def sd(g):
    return {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}


def test_oracle_matches_the_reference_fixture(g, sd):
    torch.set_num_threads(1)
    vol = torch.from_numpy(g["vol_ncdhw"])
    p = torch.from_numpy(g["p0"]).requires_grad_()
    logits = tp.convonet_decode(sd, p, {"grid": vol})
    (logits * torch.from_numpy(g["gl"])).sum().backward()
    assert np.array_equal(logits.detach().numpy(), g["logits"]) and np.array_equal(p.grad.numpy(), g["grad_p"])
    got = tp.optimize_points(lambda q: tp.convonet_decode(sd, q, {"grid": vol}), torch.from_numpy(g["p0"]), iterations=11, normalize=False)
    assert np.array_equal(got, g["trace/xyz_11"])
    volf, idx = tp.convonet_grid_features(torch.from_numpy(g["enc/feat"]), torch.from_numpy(g["enc/p"]), reso=16)
    assert np.array_equal(volf.numpy(), g["enc/vol"]) and np.array_equal(idx[:, 0].numpy(), g["enc/index"])


def test_host_instantiation_of_the_kernel_arithmetic(g, mathcheck):
    """grid_point.cuh compiled for the host (tests/mathcheck): logits and d/dp against the reference, incl. the points beyond
    the padded cube (clamps of normalize_3d_coordinate, border clipping of grid_sample)."""
    vol = np.ascontiguousarray(g["vol_ncdhw"].transpose(0, 2, 3, 4, 1))
    B, K, _ = g["p0"].shape
    lo, gr = np.zeros((B, K), np.float32), np.zeros((B, K, 3), np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    mathcheck.mc_convonet_grid_decode(P(g["dec_blob"]), P(vol), P(np.ascontiguousarray(g["p0"])), B, K, 16, 5, ctypes.c_double(0.1), 1,
                                      P(np.ascontiguousarray(g["gl"])), ctypes.c_float(0), ctypes.c_float(0), P(lo), P(gr))
    assert np.abs(lo - g["logits"]).max() < 2e-6
    assert np.abs(gr - g["grad_p"]).max() < 2e-6 * np.abs(g["grad_p"]).max()


@pytest.mark.gpu
def test_grid_kernels_through_the_c_abi(g, sd):
    from ifdefense_b200 import capi, convonet
    dec = convonet.ConvONetDecoder(sd, padding=0.1)
    assert np.array_equal(dec.blob.cpu().numpy(), g["dec_blob"])
    vol = torch.from_numpy(g["vol_ncdhw"]).cuda()
    p = torch.from_numpy(g["p0"]).cuda().requires_grad_()
    logits = dec.decode(p, {"grid": vol}).logits                        # the autograd seam (opt_defense.py:212)
    (logits * torch.from_numpy(g["gl"]).cuda()).sum().backward()
    assert np.abs(logits.detach().cpu().numpy() - g["logits"]).max() < 2e-6
    assert np.abs(p.grad.cpu().numpy() - g["grad_p"]).max() < 2e-6 * np.abs(g["grad_p"]).max()
    rest = convonet.Restorer(dec, threshold=0.2, lr=1e-3)
    for it, tol in ((0, 1e-6), (1, 1e-6), (11, 5e-6)):                   # optimize_points, raw coordinates
        P = rest.params(2, 500., it)
        P.normalize_out = 0
        x = torch.from_numpy(g["p0"]).cuda().clone()
        L = capi.lib()
        ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(2, x.shape[1]), dtype=torch.uint8, device="cuda")
        vcl = convonet.volume_to_channels_last(vol)
        capi.check(L.ifd_convonet_grid_opt(capi.ptr(vcl), capi.ptr(dec.blob), capi.ptr(x), None, None, 2, x.shape[1], 16, 32, 32, 5,
                                           ctypes.byref(P), None, capi.ptr(ws), ws.numel(), capi.stream()), "ifd_convonet_grid_opt")
        assert np.abs(x.cpu().numpy() - g["trace/xyz_%d" % it]).max() < tol, it
    out = rest.optimize_points(torch.from_numpy(g["p0"]), None, {"grid": vol}, rep_weight=500., iterations=11, printing=True)
    assert out.shape == g["p0"].shape and np.isfinite(out).all() and abs(np.linalg.norm(out, axis=2).max(1) - 1).max() < 1e-6
    assert rest.last_stats.shape == (1, 4)
    # encoder side: bins of normalize_3d_coordinate + coordinate2index('3d'), scatter_mean into the channels-last volume
    pin, feat = torch.from_numpy(g["enc/p"]).cuda(), torch.from_numpy(g["enc/feat"]).cuda().contiguous()
    B, T, C = feat.shape
    bins = torch.empty((B, T), dtype=torch.int32, device="cuda")
    capi.check(capi.lib().ifd_grid_bins(capi.ptr(pin), B, T, 16, 0.1, capi.ptr(bins), capi.stream()), "ifd_grid_bins")
    assert np.array_equal(bins.cpu().numpy(), g["enc/index"])
    outv = torch.empty((B, 16 ** 3, C), dtype=torch.float32, device="cuda")
    capi.check(capi.lib().ifd_scatter_mean_cl(capi.ptr(feat), capi.ptr(bins), B, T, C, 16 ** 3, capi.ptr(outv), capi.stream()), "ifd_scatter_mean_cl")
    want = g["enc/vol"].reshape(B, C, -1).transpose(0, 2, 1)
    assert np.array_equal(outv.cpu().numpy(), want)
