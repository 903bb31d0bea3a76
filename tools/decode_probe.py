"""Numerics probe of the decode kernels (run on a B200): the BCE gradient of kernels 2 / 5 against a float64 evaluation of
the oracle, and the loop's deviation from the reference trace of the fixture after 1 / 2 / 10 / 20 steps."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ifdefense_b200 import convonet, synth  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
from tests.gpu_util import run_opt  # noqa: E402
from tests.test_gpu_convonet import bce_grad  # noqa: E402

conv = dict(np.load(os.path.join(ROOT, "tests", "golden", "convonet.npz")))
sd = {k[3:]: torch.from_numpy(v) for k, v in conv.items() if k.startswith("sd/")}
cpl = {k: torch.from_numpy(conv["planes_nchw"][i]) for i, k in enumerate(("xz", "xy", "yz"))}
dec = convonet.ConvONetDecoder(sd, padding=0.1)
planes = convonet.planes_to_channels_last({k: v.cuda() for k, v in cpl.items()})


def decode64(sd_, p, c_plane, padding=0.1, n_blocks=5):
    """oracle.torch_port.convonet_decode with every tensor in float64 (its .float() casts dropped)."""
    c = 0
    for plane in ("xz", "xy", "yz"):
        xy = tp.normalize_coordinate(p.clone(), padding=padding, plane=plane)
        vgrid = 2.0 * xy[:, :, None] - 1.0
        c = c + F.grid_sample(c_plane[plane], vgrid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1)
    c = c.transpose(1, 2)
    net = tp._lin(sd_, "decoder.fc_p", p)
    for i in range(n_blocks):
        net = net + tp._lin(sd_, "decoder.fc_c.%d" % i, c)
        net = tp._resblock_fc(sd_, "decoder.blocks.%d" % i, net)
    return tp._lin(sd_, "decoder.fc_out", F.relu(net)).squeeze(-1)


def truth(sd_, p_, c_):
    p = p_.double().requires_grad_()
    lg = decode64({k: v.double() for k, v in sd_.items()}, p, {k: v.double() for k, v in c_.items()})
    (F.binary_cross_entropy_with_logits(lg, torch.full_like(lg, 0.2), reduction="none").mean() * p.shape[1]).backward()
    return p.grad.numpy()


want = truth(sd, torch.from_numpy(conv["p0"]), cpl)
scale = np.abs(want).max()
for k in (2, 5):
    g = bce_grad(dec, planes, conv["p0"], k)
    e = np.abs(g - want) / scale
    print("fixture grad kernel %d vs float64: max %.3e  p99 %.3e  median %.3e" % (k, e.max(), np.quantile(e, 0.99), np.median(e)))
case = synth.make_case(8, K=1024, seed=3)
want = truth(case.sd, case.p0, case.c)
scale = np.abs(want).max()
d8 = convonet.ConvONetDecoder(case.sd, padding=0.1)
p8 = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
for k in (2, 5):
    g = bce_grad(d8, p8, case.p0.numpy(), k)
    e = np.abs(g - want) / scale
    print("8 x 1024 grad kernel %d vs float64: max %.3e  p99 %.3e  median %.3e" % (k, e.max(), np.quantile(e, 0.99), np.median(e)))
for k in (2, 5):
    for n in (1, 2, 10, 20):
        x, _ = run_opt(dec, planes, conv["p0"], n, decode_kernel=k)
        d = np.abs(x - conv["trace/xyz_%d" % (n - 1)])
        print("fixture loop kernel %d steps %2d: max %.3e  p99 %.3e  n>1e-5: %d  n>1e-4: %d" % (k, n, d.max(), np.quantile(d, 0.99), (d > 1e-5).sum(), (d > 1e-4).sum()))
