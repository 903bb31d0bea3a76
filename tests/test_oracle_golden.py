"""The oracle (oracle/torch_port.py + oracle/c) against the committed fixtures generated from the real
reference (tests/golden/make_golden.py).  Index results: bit-exact.  Float results: the tolerance allows
for a different host CPU (BLAS kernel selection) than the one that generated the fixtures."""
import numpy as np
import torch

from oracle import c_oracle as co
from oracle import torch_port as tp

torch.set_num_threads(1)


def test_c_oracle_knn_bit_exact(geo):
    assert np.array_equal(co.knn(geo["xyz"], 5, 1), geo["knn5"])
    assert np.array_equal(co.knn(geo["xyz"], 20, 0), geo["dgcnn_knn20"])


def test_c_oracle_fps_ball_sor_bit_exact(geo):
    x = geo["xyz"]
    f1 = co.fps(x, 512, geo["fps_start"])
    assert np.array_equal(f1, geo["fps512"])
    new = np.take_along_axis(x, f1[..., None].astype(np.int64), 1)
    assert np.array_equal(co.ball_query(0.2, 32, x, new), geo["ball_0.2_32"])
    f2 = co.fps(new, 128, geo["fps2_start"])
    assert np.array_equal(f2, geo["fps128"])
    new2 = np.take_along_axis(new, f2[..., None].astype(np.int64), 1)
    assert np.array_equal(co.ball_query(0.4, 64, new, new2), geo["ball_0.4_64"])
    assert np.array_equal(co.fps(x, 64, geo["fps_start"]), geo["fps64_defense"])
    keep, _ = co.sor(geo["sor_xyz"], 2, 1.1)
    assert np.array_equal(keep.astype(np.uint8), geo["sor_keep"])


def test_torch_port_repulsion(geo):
    x = torch.from_numpy(geo["xyz"]).requires_grad_()
    loss = tp.repulsion_loss(x, idx=torch.from_numpy(geo["knn5"]).long())
    (loss * torch.from_numpy(geo["rep_grad_loss"])).sum().backward()
    np.testing.assert_allclose(loss.detach().numpy(), geo["rep_loss"], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), geo["rep_grad"], rtol=1e-5, atol=1e-9)


def test_torch_port_duplicate_points(geo):
    """Exact duplicates: column 0 is not necessarily self; whichever of the tied pair the
    tie order keeps, its distance is 0 -> clamped to eps -> zero gradient, same loss."""
    x = torch.from_numpy(geo["dup_xyz"]).requires_grad_()
    loss = tp.repulsion_loss(x, idx=torch.from_numpy(co.knn(geo["dup_xyz"], 5, 1)).long())
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().numpy(), geo["dup_rep_loss"], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), geo["dup_rep_grad"], rtol=1e-5, atol=1e-9)


def test_torch_port_decode(conv, conv_sd, conv_planes):
    for pk, lk, gk in (("p0", "logits", "grad_p"), ("clamp_p", "clamp_logits", "clamp_grad_p")):
        p = torch.from_numpy(conv[pk]).requires_grad_()
        logits = tp.convonet_decode(conv_sd, p, conv_planes)
        (logits * torch.from_numpy(conv["gl"])).sum().backward()
        np.testing.assert_allclose(logits.detach().numpy(), conv[lk], atol=2e-6)
        np.testing.assert_allclose(p.grad.numpy(), conv[gk], rtol=1e-4, atol=1e-5)


def test_torch_port_loop_20_steps(conv, conv_sd, conv_planes):
    got = tp.optimize_points(lambda p: tp.convonet_decode(conv_sd, p, conv_planes), torch.from_numpy(conv["p0"]),
                             rep_weight=500., iterations=19)
    assert np.abs(got - conv["final_20_normalized"]).max() < 1e-4    # north_star tolerance


def test_torch_port_encoder_matches_fixture_planes(conv, conv_sd):
    from ifdefense_b200 import models
    sd = models.synthetic_state_dict("convonet", 0)
    got = tp.convonet_encode(sd, torch.from_numpy(conv["sel"]))
    for i, k in enumerate(("xz", "xy", "yz")):
        np.testing.assert_allclose(got[k].numpy(), conv["planes_nchw"][i], atol=1e-5)
