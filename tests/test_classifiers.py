"""The consumer of the defense output: DGCNN / PointNet++ mirrors against fixtures generated from the
reference's own model classes (tests/golden/make_classifier_golden.py).  CPU: checkpoint interface (state_dict keys) and
the npz wire format; -m gpu: logits with the geometry ops on the library's kernels."""
import os

import numpy as np
import pytest
import torch

from ifdefense_b200 import classifiers, synth

from .conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "classifiers.npz"))


def test_state_dict_keys_equal_the_reference(gold):
    assert sorted(classifiers.DGCNN().state_dict().keys()) == list(gold["dgcnn_keys"])
    assert sorted(classifiers.PointNet2ClsSsg().state_dict().keys()) == list(gold["pointnet2_keys"])
    sd = {"module." + k: v for k, v in classifiers.DGCNN().state_dict().items()}       # a DataParallel checkpoint
    classifiers.DGCNN().load_state_dict(classifiers.strip_data_parallel(sd), strict=True)


def test_npz_wire_format_roundtrip(tmp_path, gold):
    """What driver.defend_npz_test_data writes (test_pc float32 [n,1024,3], labels uint8) is what load_data reads
    (baselines/dataset/ModelNet40.py:9-16)."""
    path = str(tmp_path / "convonet_opt-x.npz")
    np.savez(path, test_pc=gold["clouds"], test_label=np.array([0, 7, 20, 39], dtype=np.uint8),
             target_label=np.array([1, 2, 3, 4], dtype=np.uint8))
    pc, label = classifiers.load_npz(path)
    assert pc.dtype == np.float32 and pc.shape == (4, 1024, 3) and label.dtype == np.uint8
    assert len(classifiers.load_npz(path, "attack")) == 3
    p = gold["clouds"][1] * 3 + 1
    q = classifiers.normalize_points_np(p)
    assert abs(np.sqrt((q ** 2).sum(1)).max() - 1) < 1e-6 and np.abs(q.mean(0)).max() < 1e-6


@pytest.mark.gpu
def test_dgcnn_logits_equal_reference(gold):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = synth.fill_classifier_state_(classifiers.DGCNN(1024, 20, 40), seed=1).cuda().eval()
    x = torch.from_numpy(gold["clouds"]).cuda().transpose(1, 2).contiguous()
    from ifdefense_b200.defense import pn_utils
    assert np.array_equal(pn_utils.dgcnn_knn(x, 20).cpu().numpy(), gold["dgcnn_knn0"])      # first layer: exact indices
    with torch.no_grad():
        logits = m(x).cpu().numpy()
    err = np.abs(logits - gold["dgcnn_logits"]).max()
    print("dgcnn max |dlogit|", err, "scale", np.abs(gold["dgcnn_logits"]).max())
    # deeper layers search neighbours in feature space: features differ by fp32 rounding between cuDNN and the CPU, so a
    # near-tie may pick another neighbour; the max-pooled result moves by that rounding only
    assert err < 2e-3 * max(1.0, np.abs(gold["dgcnn_logits"]).max())
    assert np.array_equal(logits.argmax(1), gold["dgcnn_logits"].argmax(1))


@pytest.mark.gpu
def test_pointnet2_logits_equal_reference(gold):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = synth.fill_classifier_state_(classifiers.PointNet2ClsSsg(40), seed=2).cuda().eval()
    x = torch.from_numpy(gold["clouds"]).cuda().transpose(1, 2).contiguous()
    starts = (torch.from_numpy(gold["pointnet2_start1"]), torch.from_numpy(gold["pointnet2_start2"]))
    with torch.no_grad():
        logits = m(x, starts).cpu().numpy()
    err = np.abs(logits - gold["pointnet2_logits"]).max()
    print("pointnet2 max |dlogit|", err, "scale", np.abs(gold["pointnet2_logits"]).max())
    assert err < 1e-3 * max(1.0, np.abs(gold["pointnet2_logits"]).max())
    assert np.array_equal(logits.argmax(1), gold["pointnet2_logits"].argmax(1))


@pytest.mark.gpu
def test_restore_then_classify(tmp_path, gold):
    """The loop the paper's tables measure: defend an npz (ConvONet-Opt, few iterations) -> npz -> inference."""
    from ifdefense_b200 import driver, models
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    d = driver.Defender(model.cuda().eval(), driver.Args(iterations=5, batch_size=4, sor=False))
    src = str(tmp_path / "attack.npz")
    np.savez(src, test_pc=gold["clouds"], test_label=np.array([3, 3, 5, 9], dtype=np.uint8), target_label=np.array([1, 1, 1, 1], dtype=np.uint8))
    out = driver.defend_npz_test_data(d, src, rng=np.random.default_rng(0), gen=torch.Generator().manual_seed(0))
    clf = synth.fill_classifier_state_(classifiers.DGCNN(), seed=1).cuda().eval()
    acc = classifiers.test_normal(clf, out, batch_size=3)
    acc2, succ = classifiers.test_target(clf, out, batch_size=3)
    assert 0.0 <= acc <= 1.0 and acc == acc2 and 0.0 <= succ <= 1.0
    pc, label = classifiers.load_npz(out)
    assert pc.shape == (4, 1024, 3) and pc.dtype == np.float32 and label.dtype == np.uint8
