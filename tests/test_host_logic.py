"""Host-side logic that needs no GPU: weight packing, driver pre/post-processing, checkpoint shells,
sharding plan."""
import numpy as np
import torch

from ifdefense_b200 import driver, models, synth, weights
from oracle import torch_port as tp


def test_pack_convonet_decoder_layout(conv, conv_sd):
    blob = weights.pack_convonet_decoder(conv_sd)
    assert blob.dtype == np.float32 and blob.shape == (16001,)
    assert np.array_equal(blob, conv["dec_blob"])
    H = 32
    # spot-check the documented layout: fc_p.W^T, fc_p.b, then per block fc_c / fc_0 / fc_1 (W^T, b), fc_out
    assert np.array_equal(blob[:3 * H].reshape(3, H), conv_sd["decoder.fc_p.weight"].numpy().T)
    assert np.array_equal(blob[3 * H:4 * H], conv_sd["decoder.fc_p.bias"].numpy())
    layer = H * H + H
    off = 4 * H + 2 * 3 * layer + layer        # block 2, fc_0
    assert np.array_equal(blob[off:off + H * H].reshape(H, H), conv_sd["decoder.blocks.2.fc_0.weight"].numpy().T)
    assert np.array_equal(blob[-H - 1:-1], conv_sd["decoder.fc_out.weight"].numpy().reshape(-1))
    assert weights.convonet_decoder_dims(conv_sd) == (32, 32, 5)


def test_preprocess_and_init_match_oracle():
    raw = synth.clouds(3)
    for i in range(3):
        a, s = driver.preprocess_pc(raw[i], 600, 0.9, np.random.default_rng(5))
        a2, s2 = tp.preprocess_pc_np(raw[i], 600, 0.9, np.random.default_rng(5))
        assert np.array_equal(a, a2) and np.array_equal(s, s2) and s.shape == (600, 3)
        assert (a.max(0) - a.min(0)).max() <= 0.9 + 1e-6 and np.abs(a.mean(0)).max() < 1e-6
    pcs = [driver.preprocess_pc(raw[i], None, 0.9)[0][: 900 + 50 * i] for i in range(3)]      # ragged, like after SOR
    p = driver.init_points(pcs, 1024, 0.01, 0.9, torch.Generator().manual_seed(1))
    q = tp.init_points(pcs, 1024, 0.01, 0.9, torch.Generator().manual_seed(1))
    assert torch.equal(p, q) and p.shape == (3, 1024, 3) and p.abs().max() <= 0.45


def test_checkpoint_shells_have_reference_names():
    sd = models.build_convonet().state_dict()
    assert len(sd) == 99 and sum(v.numel() for k, v in sd.items() if k.startswith("decoder.")) == 16001
    assert sd["encoder.unet.up_convs.0.upconv.weight"].shape == (256, 128, 2, 2)
    assert sd["encoder.blocks.3.shortcut.weight"].shape == (32, 64)
    so = models.build_onet().state_dict()
    assert len(so) == 130 and so["decoder.block2.bn_1.conv_gamma.weight"].shape == (256, 512, 1)
    assert sum(v.numel() for k, v in so.items() if k.startswith("decoder.") and "num_batches" not in k and "running" not in k) == 3548417


def test_product_encoder_equals_oracle_encoder_cpu():
    case = synth.make_case(2, K=64, seed=0)
    got = tp.convonet_encode(case.sd, case.sel)
    for k in got:
        assert torch.equal(got[k], case.c[k])


def test_npz_wire_format(tmp_path):
    """opt_defense.py:317-344: keys, dtypes and the output path convention."""
    class Fake:
        def defend_point_cloud(self, pc, **kw):
            return np.zeros((len(pc), 1024, 3), np.float64) + 0.25
    f = tmp_path / "adv.npz"
    np.savez(f, test_pc=np.zeros((3, 1024, 6), np.float32), test_label=np.arange(3), target_label=np.arange(3) + 1)
    out = driver.defend_npz_test_data(Fake(), str(f))
    assert out.endswith("ConvONet-Opt/convonet_opt-adv.npz")
    z = np.load(out)
    assert z["test_pc"].dtype == np.float32 and z["test_pc"].shape == (3, 1024, 3)
    assert z["test_label"].dtype == np.uint8 and z["target_label"].dtype == np.uint8


def test_mesh_and_classifier_paths_fail_loudly_without_gpu():
    """No CPU fallback: the mesh tail and the classifier geometry ops raise when there is no CUDA device."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from ifdefense_b200 import classifiers, mesh
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mesh.marching_cubes(np.zeros((4, 4, 4)), 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mesh.MISE(4, 1, 0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        classifiers.DGCNN()(torch.zeros(1, 3, 64))


def test_encode_chunked_places_clouds_at_job_aligned_positions():
    """Defender.encode_chunked (host logic, CPU stub): every encoder call sees exactly `encoder_chunk` clouds, cloud i of the
    job sits at position i % chunk, and the outputs come back in slice order."""
    import types
    import torch
    from ifdefense_b200 import driver
    calls = []

    class Model:
        def encode_inputs(self, x):
            calls.append(x.clone())
            pos = torch.arange(x.shape[0], dtype=torch.float32).view(-1, 1)
            return {"xz": torch.cat([x[:, 0, :1], pos], dim=1)}          # (cloud tag, position in the batch)

    stub = types.SimpleNamespace(args=types.SimpleNamespace(encoder_chunk=4), model=Model())
    sel = torch.arange(100, 107, dtype=torch.float32).view(7, 1, 1).expand(7, 5, 3).contiguous()    # clouds tagged 100..106
    for first in (0, 1, 3, 4, 6):
        calls.clear()
        out = driver.Defender.encode_chunked(stub, sel, first)["xz"]
        assert out.shape == (7, 2)
        assert torch.equal(out[:, 0], torch.arange(100, 107, dtype=torch.float32))                   # slice order kept
        assert torch.equal(out[:, 1], ((torch.arange(7) + first) % 4).float())                       # job-aligned position
        assert all(c.shape[0] == 4 for c in calls)                                                   # always full chunks
        assert len(calls) == (first % 4 + 7 + 3) // 4


def test_restorer_part_policy():
    from ifdefense_b200 import convonet
    r = convonet.Restorer.__new__(convonet.Restorer)
    r.side_by_side = True
    # batches of >= 96 clouds run as equal parts of at most 64 clouds side by side (64 clouds = one launch of the one-CTA tail on
    # 64 of the 148 SMs; the library keeps up to four loops in flight); a prime count stays one loop
    assert [r._parts(b) for b in (1, 64, 95, 96, 127, 128, 164, 192, 193, 256, 384)] == [1, 1, 1, 2, 1, 2, 4, 3, 1, 4, 6]
    r.side_by_side = False
    assert r._parts(192) == 1
    r.side_by_side = 4
    assert r._parts(192) == 4 and r._parts(190) == 1


def test_bench_emits_exactly_one_stdout_line_despite_library_chatter():
    """bench.py's contract is one JSON line on stdout; NCCL prints its version banner to fd 1 under torchrun.  After
    _protect_stdout() everything written to fd 1 goes to stderr and emit() alone reaches the real stdout."""
    import subprocess
    import sys
    from .conftest import ROOT
    code = ("import bench, os\nbench._protect_stdout()\nos.write(1, b'NCCL version x\\n')\nprint('chatter')\n"
            "bench.emit('{\"a\": 1}')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and r.stdout == '{"a": 1}\n'
    assert "NCCL version x" in r.stderr and "chatter" in r.stderr
