"""Generate tests/golden/grid.npz from the UNMODIFIED reference (/root/reference) on CPU: the 'grid' (feature volume,
trilinear) variant of ConvONet -- LocalDecoder.sample_grid_feature (ConvONet/src/conv_onet/models/decoder.py:59-67) and
LocalPoolPointnet.generate_grid_features (ConvONet/src/encoder/pointnet.py:88-99).  No shipped config selects this variant
(SURVEY.md F2), so the classes are built directly: c_dim = hidden = 32, 5 blocks, a 16^3 volume.

Run in the build container only:   python tests/golden/make_grid_golden.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import torch_port as tp  # noqa: E402
from ifdefense_b200 import weights  # noqa: E402

torch.set_num_threads(1)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grid.npz")


def main():
    ns = ref_import.load("ConvONet")
    torch.manual_seed(5)
    R, B, K, T, C = 16, 2, 128, 96, 32
    dec = ns.decoder.LocalDecoder(dim=3, c_dim=C, hidden_size=32, n_blocks=5, padding=0.1).eval()
    for m in dec.modules():
        if hasattr(m, "fc_1"):                                  # the reference zero-initialises fc_1 (layers.py:37)
            torch.nn.init.kaiming_uniform_(m.fc_1.weight, a=5 ** 0.5)
    for q in dec.parameters():
        q.requires_grad = False
    sd = {"decoder." + k: v for k, v in dec.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    vol = torch.randn(B, C, R, R, R, generator=g) * 0.5
    p0 = (torch.rand(B, K, 3, generator=g) - 0.5) * 0.9
    p0[:, :6] = (torch.rand(B, 6, 3, generator=g) - 0.5) * 1.3   # beyond the padded cube: both clamps, border clipping
    gl = torch.randn(B, K, generator=g)
    out = {"vol_ncdhw": vol.numpy(), "p0": p0.numpy(), "gl": gl.numpy(), "dec_blob": weights.pack_convonet_decoder(sd, "decoder.")}
    out.update({"sd/" + k: v.numpy() for k, v in sd.items()})
    p = p0.clone().requires_grad_()
    logits = dec(p, {"grid": vol})
    (logits * gl).sum().backward()
    out["logits"], out["grad_p"] = logits.detach().numpy(), p.grad.numpy()
    assert torch.equal(logits.detach(), tp.convonet_decode(sd, p0, {"grid": vol}))
    # 12 steps of the restated optimize_points around the reference's decoder and RepulsionLoss (opt_defense.py:182-239)
    pts = p0.clone().requires_grad_()
    opt = torch.optim.Adam([pts], lr=1e-3)
    thr = torch.ones(B, K) * 0.2
    trace = {}
    for i in range(12):
        occ = dec(pts, {"grid": vol})
        loss = F.binary_cross_entropy_with_logits(occ, thr, reduction="none").mean() * K + ns.repulsion.repulsion_loss(pts).mean() * 500.
        opt.zero_grad()
        loss.backward()
        opt.step()
        if i in (0, 1, 11):
            trace[i] = pts.detach().clone().numpy()
    for i, v in trace.items():
        out["trace/xyz_%d" % i] = v
    # encoder side: scatter_mean of point features into the volume
    enc = ns.encoder.LocalPoolPointnet(c_dim=C, dim=3, hidden_dim=32, scatter_type="max", unet=False, unet3d=False,
                                       grid_resolution=R, plane_type=["grid"], padding=0.1, n_blocks=5)
    pin = (torch.rand(B, T, 3, generator=g) - 0.5) * 1.15
    feat = torch.randn(B, T, C, generator=g)
    volf = enc.generate_grid_features(pin, feat)
    idx = ns.common.coordinate2index(ns.common.normalize_3d_coordinate(pin.clone(), padding=0.1), R, coord_type="3d")
    out.update({"enc/p": pin.numpy(), "enc/feat": feat.numpy(), "enc/vol": volf.numpy(), "enc/index": idx[:, 0].numpy().astype(np.int32)})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
