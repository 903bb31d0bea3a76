"""-m gpu: the geometry kernels through the C ABI against the golden fixtures (generated from the real
reference) and the C oracle.  Index results are compared bit-exactly."""
import numpy as np
import pytest
import torch

from ifdefense_b200 import capi, defense
from oracle import c_oracle as co
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


def rand_clouds(B, N, C=3, seed=0):
    return np.random.default_rng(seed).uniform(-0.45, 0.45, size=(B, N, C)).astype(np.float32)


def test_knn_golden_bit_exact(geo):
    idx = defense.knn_point(5, dev(geo["xyz"]))
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), geo["knn5"])
    idx20 = defense.dgcnn_knn(dev(geo["xyz"]).transpose(2, 1), 20)
    assert np.array_equal(idx20.cpu().numpy(), geo["dgcnn_knn20"])


@pytest.mark.parametrize("B,N,k,drop", [(8, 1024, 5, 1), (3, 1000, 5, 1), (2, 37, 5, 1), (1, 4096, 5, 1), (2, 2048, 20, 0),
                                        (2, 300, 2, 1), (1, 6, 5, 1), (2, 513, 31, 1)])
def test_knn_vs_c_oracle(B, N, k, drop):
    x = rand_clouds(B, N, seed=N + k)
    want, wkeys = co.knn(x, k, drop, want_keys=True)
    xd = dev(x)
    idx = torch.empty((B, N, k), dtype=torch.int32, device="cuda")
    keys = torch.empty((B, N, k), dtype=torch.float32, device="cuda")
    capi.check(capi.lib().ifd_knn(capi.ptr(xd), B, N, 3, k, drop, capi.ptr(idx), capi.ptr(keys), capi.stream()))
    assert np.array_equal(idx.cpu().numpy(), want)
    assert np.array_equal(keys.cpu().numpy(), wkeys)            # the ranked distances themselves are bit-equal


@pytest.mark.parametrize("C", [4, 64, 128])
def test_knn_feature_space(C):
    """DGCNN layers 2-4 rank in feature space; the kernel and the C oracle share one FMA-chain definition."""
    x = (np.random.default_rng(C).normal(size=(2, 512, C)) * 0.5).astype(np.float32)
    want = co.knn(x, 20, 0)
    got = defense.dgcnn_knn(dev(x).transpose(2, 1), 20)
    assert np.array_equal(got.cpu().numpy(), want)


def test_knn_ties_lowest_index_first():
    """Exact duplicates: the tie policy is lowest index first; torch.topk's is unspecified."""
    x = rand_clouds(1, 64, seed=1)
    x[0, 1] = x[0, 0]
    x[0, 9] = x[0, 7]
    got = defense.knn_point(5, dev(x)).cpu().numpy()
    assert np.array_equal(got, co.knn(x, 5, 1))
    assert got[0, 0, 0] == 1 and got[0, 1, 0] == 1      # rows 0 and 1 both drop index 0 and keep index 1


def test_knn_errors():
    x = dev(rand_clouds(1, 4))
    with pytest.raises(RuntimeError, match="exceeds N"):
        defense.knn_point(5, x)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        defense.dgcnn_knn(dev(rand_clouds(1, 64, C=5)).transpose(2, 1), 4)


def test_repulsion_golden(geo):
    x = dev(geo["xyz"]).requires_grad_()
    loss = defense.repulsion_loss(x)                                   # nn.Module seam, [B]
    (loss * dev(geo["rep_grad_loss"])).sum().backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), geo["rep_loss"], rtol=2e-6)
    g = x.grad.cpu().numpy()
    assert np.abs(g - geo["rep_grad"]).max() < 2e-6 * np.abs(geo["rep_grad"]).max()
    # idx through the raw entry point
    B, K, _ = geo["xyz"].shape
    L = capi.lib()
    ws = torch.empty(L.ifd_knn_repulsion_workspace_bytes(B, K), dtype=torch.uint8, device="cuda")
    idx = torch.empty((B, K, 5), dtype=torch.int32, device="cuda")
    capi.check(L.ifd_knn_repulsion(capi.ptr(x.detach()), B, K, 5, 0.07, 0.03, 1e-12, None, capi.ptr(idx), None, None,
                                   capi.ptr(ws), ws.numel(), capi.stream()))
    assert np.array_equal(idx.cpu().numpy(), geo["knn5"])


def test_repulsion_duplicates_and_determinism(geo):
    x = dev(geo["dup_xyz"]).requires_grad_()
    loss = defense.repulsion_loss(x)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), geo["dup_rep_loss"], rtol=2e-6)
    assert np.abs(x.grad.cpu().numpy() - geo["dup_rep_grad"]).max() < 2e-6 * np.abs(geo["dup_rep_grad"]).max()
    big = dev(rand_clouds(16, 1024, seed=3) * 0.2)          # dense: many scatter collisions per point
    outs = []
    for _ in range(3):
        b = big.clone().requires_grad_()
        defense.repulsion_loss(b).sum().backward()
        outs.append(b.grad.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])   # bitwise reproducible scatter-add


def test_fps_golden(geo):
    x = dev(geo["xyz"])
    f1 = defense.farthest_point_sample(x, 512, torch.from_numpy(geo["fps_start"]).long())
    assert np.array_equal(f1.cpu().numpy(), geo["fps512"])
    new = defense.index_points(x, f1)
    f2 = defense.farthest_point_sample(new, 128, torch.from_numpy(geo["fps2_start"]).long())
    assert np.array_equal(f2.cpu().numpy(), geo["fps128"])
    f3 = defense.farthest_point_sample(x, 64, torch.from_numpy(geo["fps_start"]).long())
    assert np.array_equal(f3.cpu().numpy(), geo["fps64_defense"])


@pytest.mark.parametrize("N,npoint", [(1024, 1024), (4096, 512), (100, 7), (8192, 64)])
def test_fps_vs_c_oracle(N, npoint):
    x = rand_clouds(3, N, seed=N)
    x[1, N // 2:] = x[1, : N - N // 2]                     # duplicated half: equal maxima -> first occurrence
    st = np.array([0, N - 1, N // 3], np.int32)
    got = defense.farthest_point_sample(dev(x), npoint, torch.from_numpy(st).long())
    assert np.array_equal(got.cpu().numpy(), co.fps(x, npoint, st))


def test_ball_query_golden(geo):
    x = dev(geo["xyz"])
    new = defense.index_points(x, torch.from_numpy(geo["fps512"]).long().cuda())
    assert np.array_equal(defense.query_ball_point(0.2, 32, x, new).cpu().numpy(), geo["ball_0.2_32"])
    new2 = defense.index_points(new, torch.from_numpy(geo["fps128"]).long().cuda())
    assert np.array_equal(defense.query_ball_point(0.4, 64, new, new2).cpu().numpy(), geo["ball_0.4_64"])


def test_ball_query_vs_c_oracle_with_empty_balls():
    x = rand_clouds(2, 3000, seed=9)
    centres = rand_clouds(2, 200, seed=10) * 3.0            # most centres are far outside: no hit -> N
    want = co.ball_query(0.1, 16, x, centres)
    got = defense.query_ball_point(0.1, 16, dev(x), dev(centres)).cpu().numpy()
    assert np.array_equal(got, want) and (got == 3000).any()


def test_sor_golden(geo):
    kept = defense.SORDefense(k=2, alpha=1.1)(dev(geo["sor_xyz"]))
    assert [len(k) for k in kept] == geo["sor_keep"].sum(1).tolist()            # ragged output
    for b in range(4):
        assert np.array_equal(kept[b].cpu().numpy(), geo["sor_xyz"][b][geo["sor_keep"][b].astype(bool)])
    # the float64 statistic itself
    B, K, _ = geo["sor_xyz"].shape
    keep = torch.empty((B, K), dtype=torch.uint8, device="cuda")
    val = torch.empty((B, K), dtype=torch.float64, device="cuda")
    capi.check(capi.lib().ifd_sor(capi.ptr(dev(geo["sor_xyz"])), B, K, 2, 1.1, capi.ptr(keep), capi.ptr(val), capi.stream()))
    _, want = co.sor(geo["sor_xyz"], 2, 1.1)
    np.testing.assert_allclose(val.cpu().numpy(), want, rtol=1e-13, atol=1e-18)
