"""-m gpu: the ONet twin of the defense driver (ONet/opt_defense.py: SOR -> preprocess -> ResnetPointnet encoder -> init ->
ONet-Opt loop -> normalise) through driver.ONetDefender, and its npz naming.  (The loop itself is pinned in
tests/test_gpu_onet.py; independence from the split across ranks is asserted for the ConvONet driver only so far.)"""
import numpy as np
import pytest
import torch

from ifdefense_b200 import driver, models, synth

pytestmark = pytest.mark.gpu


def test_onet_defender_end_to_end(tmp_path):
    model = models.build_onet()
    model.load_state_dict(models.synthetic_state_dict("onet", 0))
    d = driver.ONetDefender(model, driver.Args(batch_size=2, iterations=3, input_npoint=300))
    pc = synth.clouds(3)                                         # batches of 2 + 1
    out = d.defend_point_cloud(pc, rng=np.random.default_rng(1), gen=torch.Generator().manual_seed(1))
    assert out.shape == (3, 1024, 3) and out.dtype == np.float32 and np.isfinite(out).all()
    assert np.abs(out.mean(1)).max() < 1e-5                      # normalize_batch_pc (opt_defense.py:76-83)
    np.testing.assert_allclose(np.linalg.norm(out, axis=2).max(1), 1.0, rtol=1e-6)
    again = d.defend_point_cloud(pc, rng=np.random.default_rng(1), gen=torch.Generator().manual_seed(1))
    assert np.array_equal(out, again)                            # reproducible run to run
    part = d.restore_slice(pc, 0, 2, 2, 4)                       # the sharded entry point runs for ONet as well
    assert part.shape == (2, 1024, 3) and np.isfinite(part).all()
    f = tmp_path / "adv.npz"
    np.savez(f, test_pc=pc, test_label=np.arange(3))
    saved = driver.defend_npz_test_data(d, str(f), rng=np.random.default_rng(0))
    assert saved.endswith("ONet-Opt/onet_opt-adv.npz") and np.load(saved)["test_pc"].shape == (3, 1024, 3)
