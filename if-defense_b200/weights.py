"""Reference checkpoints -> kernel parameter blobs.

The checkpoint interface is the reference's `state_dict`.  The kernels want every
Linear transposed to [in][out] and concatenated in the order documented in include/ifd_b200.h.
"""
import numpy as np
import torch


def _np(t):
    return t.detach().cpu().float().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float32)


def convonet_decoder_dims(sd, prefix="decoder."):
    """(C, H, n_blocks) from the state_dict of ConvolutionalOccupancyNetwork (models/decoder.py:22-48)."""
    H = sd[prefix + "fc_p.weight"].shape[0]
    C = sd[prefix + "fc_c.0.weight"].shape[1]
    n_blocks = 0
    while (prefix + "blocks.%d.fc_0.weight" % n_blocks) in sd:
        n_blocks += 1
    return int(C), int(H), n_blocks


def pack_convonet_decoder(sd, prefix="decoder."):
    """-> float32 numpy [ifd_convonet_decoder_nfloats(C, H, n_blocks)] in kernel layout."""
    C, H, nb = convonet_decoder_dims(sd, prefix)
    if C != H:
        raise RuntimeError("ConvONet decoder kernels need c_dim == hidden_size (got %d, %d)" % (C, H))
    parts = [_np(sd[prefix + "fc_p.weight"]).T, _np(sd[prefix + "fc_p.bias"])]
    for i in range(nb):
        for name in ("fc_c.%d" % i, "blocks.%d.fc_0" % i, "blocks.%d.fc_1" % i):
            parts.append(_np(sd[prefix + name + ".weight"]).T)
            parts.append(_np(sd[prefix + name + ".bias"]))
        if (prefix + "blocks.%d.shortcut.weight" % i) in sd:
            raise RuntimeError("decoder ResnetBlockFC with a shortcut is not supported")
    parts += [_np(sd[prefix + "fc_out.weight"]).reshape(-1), _np(sd[prefix + "fc_out.bias"]).reshape(-1)]
    blob = np.concatenate([np.ascontiguousarray(p, dtype=np.float32).reshape(-1) for p in parts])
    assert blob.size == 4 * H + nb * 3 * (H * H + H) + H + 1
    return blob


def planes_to_channels_last_np(c_plane):
    """dict {'xz','xy','yz': [B,C,R,R]} -> float32 numpy [3,B,R,R,C] (test helper; the product converts on device)."""
    return np.stack([_np(c_plane[k]).transpose(0, 2, 3, 1) for k in ("xz", "xy", "yz")]).copy()


ONET_CBN_ORDER = ["block%d.bn_%d" % (i, j) for i in range(5) for j in range(2)] + ["bn"]
ONET_FC_ORDER = ["block%d.fc_%d" % (i, j) for i in range(5) for j in range(2)]


def pack_onet_decoder(sd, prefix="decoder."):
    """DecoderCBatchNorm state_dict (ONet/im2mesh/onet/models/decoder.py:89-108) -> float32 blob in the order
    documented at ifd_onet_decoder_nfloats (include/ifd_b200.h).  Conv1d kernels [out][in][1] are flattened."""
    if (prefix + "fc_z.weight") in sd:
        raise RuntimeError("ONet decoder with z_dim != 0 is not supported (configs/onet_mn40.yaml has z_dim: 0)")
    parts = [_np(sd[prefix + "fc_p.weight"]).reshape(256, 3), _np(sd[prefix + "fc_p.bias"])]
    for name in ONET_CBN_ORDER:
        p = prefix + name
        parts += [_np(sd[p + ".conv_gamma.weight"]).reshape(256, 512), _np(sd[p + ".conv_gamma.bias"]),
                  _np(sd[p + ".conv_beta.weight"]).reshape(256, 512), _np(sd[p + ".conv_beta.bias"]),
                  _np(sd[p + ".bn.running_mean"]), _np(sd[p + ".bn.running_var"])]
    for name in ONET_FC_ORDER:
        parts += [_np(sd[prefix + name + ".weight"]).reshape(256, 256), _np(sd[prefix + name + ".bias"])]
    parts += [_np(sd[prefix + "fc_out.weight"]).reshape(-1), _np(sd[prefix + "fc_out.bias"]).reshape(-1)]
    blob = np.concatenate([np.ascontiguousarray(x, dtype=np.float32).reshape(-1) for x in parts])
    assert blob.size == 256 * 4 + 11 * (2 * (256 * 512 + 256) + 512) + 10 * (256 * 256 + 256) + 257
    return blob
