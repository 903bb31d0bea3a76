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


def test_preprocess_kernel_equals_numpy():
    """ifd_preprocess_pc against the numpy statements of preprocess_pc (opt_defense.py:114-130), bit for bit, with and
    without a SOR mask (ragged counts), including K that is not a multiple of the block size."""
    import torch
    from ifdefense_b200 import capi
    rng = np.random.default_rng(1)
    for K in (1024, 700, 257, 5):
        pc = (rng.standard_normal((6, K, 3)) * rng.uniform(0.2, 2.0, size=(6, 1, 3)) + rng.uniform(-1, 1, size=(6, 1, 3))).astype(np.float32)
        for masked in (False, True):
            keep = (rng.random((6, K)) < 0.8).astype(np.uint8) if masked else None
            if masked:
                keep[:, 0] = 1
            x = torch.from_numpy(pc).cuda()
            out = torch.full_like(x, 7.0)
            counts = torch.empty(6, dtype=torch.int32, device="cuda")
            kd = torch.from_numpy(keep).cuda() if masked else None
            capi.check(capi.lib().ifd_preprocess_pc(capi.ptr(x), capi.ptr(kd), 6, K, 0.9, capi.ptr(out), capi.ptr(counts), capi.stream()))
            out, counts = out.cpu().numpy(), counts.cpu().numpy()
            for b in range(6):
                p = pc[b][keep[b] != 0] if masked else pc[b]
                want, _ = driver.preprocess_pc(p, None, 0.9)
                assert counts[b] == len(p)
                assert np.array_equal(out[b, :len(p)], want), (K, masked, b)
                assert (out[b, len(p):] == 0).all()


def test_device_preprocess_path_equals_host_path():
    """defend_point_cloud with SOR + preprocess + init gathers on the device gives the bits of the per-cloud numpy path
    (the reference's own sequence) for the same RNG streams."""
    import torch
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    pc = synth.clouds(5)
    pc[0, :7] += 0.8                                         # a few outliers for SOR to remove
    outs = []
    for dev_path in (True, False):
        d = driver.Defender(model, driver.Args(batch_size=3, iterations=6, device_preprocess=dev_path))
        outs.append(d.defend_point_cloud(pc, rng=np.random.default_rng(5), gen=torch.Generator().manual_seed(5)))
    assert np.isfinite(outs[0]).all() and np.array_equal(outs[0], outs[1])


def test_sharded_slices_device_path_and_encoder_chunks():
    """restore_slice: the device pre-processing path equals the per-cloud numpy path, and with fixed-size encoder chunks
    (last one padded) a cloud's result does not depend on where the slice boundaries fall."""
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    pc = synth.clouds(6)
    pc[2, :5] -= 0.7
    dev_d = driver.Defender(model, driver.Args(batch_size=6, iterations=5, encoder_chunk=4, device_preprocess=True))
    host_d = driver.Defender(model, driver.Args(batch_size=6, iterations=5, encoder_chunk=4, device_preprocess=False))
    full = dev_d.restore_slice(pc, 0, 6, 6, 11)
    assert np.array_equal(full, host_d.restore_slice(pc, 0, 6, 6, 11))
    for cuts in ([0, 3, 6], [0, 1, 6], [0, 2, 4, 5, 6]):
        out = np.concatenate([dev_d.restore_slice(pc, a, b, 6, 11) for a, b in zip(cuts[:-1], cuts[1:])])
        assert np.array_equal(out, full), cuts
    # the encoder alone: planes of a cloud do not depend on the slice it arrives in (chunks aligned to job indices)
    import torch
    sel = torch.rand(7, 600, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(0)) - 0.5
    whole = dev_d.encode_chunked(sel, 0)
    for lo in (1, 3, 4, 6):
        part = dev_d.encode_chunked(sel[lo:], lo)
        assert all(torch.equal(whole[k][lo:], part[k]) for k in whole), lo


def test_small_clouds_use_all_points_like_the_reference():
    """K <= input_npoint with --sor=False: preprocess_pc (opt_defense.py:134-141) hands ALL points to the encoder and draws
    nothing; the device path does the same (it used to raise) and equals the per-cloud numpy path bit for bit."""
    import torch
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    pc = synth.clouds(3)[:, :500]
    outs = []
    for dev_path in (True, False):
        d = driver.Defender(model, driver.Args(batch_size=3, iterations=4, sor=False, device_preprocess=dev_path))
        outs.append(d.defend_point_cloud(pc, rng=np.random.default_rng(2), gen=torch.Generator().manual_seed(2)))
    assert outs[0].shape == (3, 1024, 3) and np.isfinite(outs[0]).all() and np.array_equal(outs[0], outs[1])
