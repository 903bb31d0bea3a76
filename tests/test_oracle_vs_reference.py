"""Pins oracle/torch_port.py bit-for-bit against the UNMODIFIED reference classes imported from
/root/reference.  Only runs where the reference tree exists (the build container); the GPU box relies on the
committed fixtures instead."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import torch_port as tp

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
torch.set_num_threads(1)


@pytest.fixture(scope="module")
def conv_ref():
    from ifdefense_b200 import synth
    ns = ref_import.load("ConvONet")
    _, model = ref_import.build_model(ns)
    case = synth.make_case(2, K=128, seed=3)
    model.load_state_dict(case.sd, strict=True)          # checkpoint interface: names and shapes
    return ns, model, case


def test_state_dict_interface(conv_ref):
    from ifdefense_b200 import models
    _, model, _ = conv_ref
    ours = models.build_convonet().state_dict()
    theirs = model.state_dict()
    assert list(ours.keys()) == list(theirs.keys()) and len(ours) == 99
    assert all(ours[k].shape == theirs[k].shape for k in ours)
    ns2 = ref_import.load("ONet")
    _, m2 = ref_import.build_model(ns2)
    o2 = models.build_onet().state_dict()
    assert list(o2.keys()) == list(m2.state_dict().keys()) and len(o2) == 130
    ref_import.load("ConvONet")


def test_encoder_decoder_bitwise(conv_ref):
    ns, model, case = conv_ref
    with torch.no_grad():
        c = model.encode_inputs(case.sel)
        c2 = tp.convonet_encode(case.sd, case.sel)
    for k in c:
        assert torch.equal(c[k], c2[k]) and torch.equal(c[k], case.c[k])
    p = case.p0.clone().requires_grad_()
    a = model.decode(p, c).logits
    b = tp.convonet_decode(case.sd, p, c)
    assert torch.equal(a, b)


def test_geometry_bitwise(conv_ref):
    ns, _, case = conv_ref
    x = case.p0
    assert torch.equal(ns.pn_utils.knn_point(5, x), tp.knn_point(5, x))
    assert torch.equal(ns.repulsion.repulsion_loss(x), tp.repulsion_loss(x))
    kept = ns.sor.SORDefense()(torch.from_numpy(case.raw))
    kept2, _, _ = tp.sor_outlier_removal(torch.from_numpy(case.raw))
    assert all(torch.equal(a, b) for a, b in zip(kept, kept2))


def test_loop_bitwise_single_thread(conv_ref):
    """Restated optimize_points around the reference's own decode/repulsion objects == torch_port loop."""
    import torch.nn.functional as F
    ns, model, case = conv_ref
    with torch.no_grad():
        c = model.encode_inputs(case.sel)
    pts = case.p0.clone().requires_grad_()
    opt = torch.optim.Adam([pts], lr=0.001)
    thr = torch.ones(pts.shape[:2]) * 0.2
    for _ in range(31):
        occ = model.decode(pts, c).logits
        loss = F.binary_cross_entropy_with_logits(occ, thr, reduction='none').mean() * pts.shape[1] + \
            ns.repulsion.repulsion_loss(pts).mean() * 500.
        opt.zero_grad()
        loss.backward()
        opt.step()
    want = pts.detach().numpy()
    got = tp.optimize_points(lambda p: tp.convonet_decode(case.sd, p, c), case.p0, iterations=30, normalize=False)
    assert np.array_equal(got, want)
