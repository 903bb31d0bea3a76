"""-m gpu: the library form of the defense driver (SOR -> preprocess -> encode -> init -> optimize -> normalise) end to
end on one GPU, and the sharding contract: a job cut into per-rank slices gives the same bits as the unsplit job."""
import numpy as np
import pytest

from ifdefense_b200 import driver, models, shard, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def defender():
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    return driver.Defender(model, driver.Args(batch_size=4, iterations=12))


def test_defender_sharded_equals_unsplit(defender):
    pc = synth.clouds(6)                                    # [6,1024,3]; batches of 4 + 2
    full = defender.defend_point_cloud_sharded(pc, seed=3, rank=0, world=1)
    assert full.shape == (6, 1024, 3) and full.dtype == np.float32 and np.isfinite(full).all()
    assert np.abs(full.mean(1)).max() < 1e-5
    np.testing.assert_allclose(np.linalg.norm(full, axis=2).max(1), 1.0, rtol=1e-6)
    for world in (2, 3):                                    # what ranks 0..world-1 would compute, assembled by hand
        out = np.zeros_like(full)
        for r, segs in enumerate(shard.plan(6, 4, world)):
            for a, b, B_ref in segs:
                out[a:b] = defender.restore_slice(pc, a, b, B_ref, 3)
        assert np.array_equal(out, full), world


def test_defender_reference_loop_runs(defender, tmp_path):
    """defend_npz_test_data with the reference's batching and global RNG semantics (opt_defense.py:255-344)."""
    pc = synth.clouds(3)
    f = tmp_path / "adv.npz"
    np.savez(f, test_pc=pc, test_label=np.arange(3), target_label=np.arange(3) + 1)
    out = driver.defend_npz_test_data(defender, str(f), rng=np.random.default_rng(0))
    z = np.load(out)
    assert z["test_pc"].shape == (3, 1024, 3) and z["test_pc"].dtype == np.float32 and np.isfinite(z["test_pc"]).all()
    assert z["test_label"].dtype == np.uint8 and z["target_label"].dtype == np.uint8
