"""ONet: oracle (oracle/torch_port.py) and host logic against the fixtures generated from the reference."""
import numpy as np
import pytest
import torch

from ifdefense_b200 import models, weights
from oracle import torch_port as tp

torch.set_num_threads(1)


@pytest.fixture(scope="module")
def onet(conftest_golden_dir):
    return dict(np.load(conftest_golden_dir + "/onet.npz"))


@pytest.fixture(scope="module")
def sd():
    return models.synthetic_state_dict("onet", 0)


def test_synthetic_weights_are_the_fixture_weights(onet, sd):
    blob = weights.pack_onet_decoder(sd)
    assert blob.size == 3548417 + 11 * 512            # parameters + the 11 BN running_mean / running_var pairs
    got = np.array([np.sum(blob.astype(np.float64)), np.sum(np.abs(blob).astype(np.float64)), blob[12345]])
    np.testing.assert_allclose(got, onet["weights_checksum"], rtol=1e-12)


def test_oracle_decode_and_encoder(onet, sd):
    for tag in ("b2", "cfg0"):
        c = tp.onet_encode(sd, torch.from_numpy(onet[tag + "/sel"]))
        np.testing.assert_allclose(c.numpy(), onet[tag + "/c"], atol=2e-5)
        p = torch.from_numpy(onet[tag + "/p0"]).requires_grad_()
        logits = tp.onet_decode(sd, p, torch.from_numpy(onet[tag + "/c"]))
        (logits * torch.from_numpy(onet[tag + "/gl"])).sum().backward()
        np.testing.assert_allclose(logits.detach().numpy(), onet[tag + "/logits"], atol=5e-6)
        assert np.abs(p.grad.numpy() - onet[tag + "/grad_p"]).max() < 1e-5 * np.abs(onet[tag + "/grad_p"]).max()


def test_oracle_config0_onet_opt_cpu(onet, sd):
    """BASELINE.json configs[0]: ONet-Opt on 1 synthetic 1024-point cloud, 20 iterations, PyTorch CPU."""
    c = torch.from_numpy(onet["cfg0/c"])
    out = tp.optimize_points(lambda p: tp.onet_decode(sd, p, c), torch.from_numpy(onet["cfg0/p0"]), rep_weight=500., iterations=20)
    assert out.shape == (1, 1024, 3) and np.isfinite(out).all()
    np.testing.assert_allclose(np.linalg.norm(out, axis=2).max(), 1.0, rtol=1e-6)
    d = np.abs(out - onet["cfg0/final_normalized"])
    assert np.median(d) < 1e-6 and (d < 1e-4).mean() > 0.99
