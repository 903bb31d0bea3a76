"""Checkpoint-compatible model shells.

The reference's checkpoint interface is the state_dict of its nn.Modules (99 tensors
for ConvONet, 130 for ONet).  These classes reproduce exactly those parameter names and shapes so that a
reference checkpoint loads with strict=True, and route the hot ops to the sm_100a kernels:

  ConvolutionalOccupancyNetwork.decode   -> convonet.ConvONetDecoder (fused gather + ResNet-MLP kernels)
  LocalPoolPointnet.forward              -> torch/cuDNN for the Linear / U-Net layers (once per batch, phase 1
                                            of SURVEY.md 7.1 step 7), reference op order otherwise
  OccupancyNetwork / ResnetPointnet      -> ONet shells (decoder kernels: see onet.py when present)

Reference classes mirrored: ConvONet/src/conv_onet/models/__init__.py:14-88, models/decoder.py:8-95,
src/encoder/pointnet.py:11-168, src/encoder/unet.py:48-239, src/layers.py:6-48;
ONet/im2mesh/onet/models/__init__.py, models/decoder.py:75-133, im2mesh/encoder/pointnet.py:60-113,
im2mesh/layers.py:58-107,186-242.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import distributions as dist

# ------------------------------------------------------------------------------------------ shared blocks


class ResnetBlockFC(nn.Module):
    def __init__(self, size_in, size_out=None, size_h=None):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)

    def forward(self, x):
        dx = self.fc_1(F.relu(self.fc_0(F.relu(x))))
        return (x if self.shortcut is None else self.shortcut(x)) + dx


class _PackedCache:
    """Weight images of a module for the tensor-core GEMM engine (tc.py), rebuilt when a parameter changes or moves."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, module, build):
        ps = list(module.parameters())
        key = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)
        if self.key != key:
            with torch.no_grad():
                self.val = build()
            self.key = key
        return self.val


def _pack_resblock(tc, blk, seg_widths):
    """ResnetBlockFC on the GEMM engine = two launches: h = fc_0(relu(x)); out = [fc_1 | shortcut] . [relu(h) | x] + b1
    (the shortcut GEMM rides in the K dimension of the second one; x itself may be a concatenation)."""
    size_h = blk.fc_0.out_features
    w1 = blk.fc_1.weight if blk.shortcut is None else torch.cat([blk.fc_1.weight, blk.shortcut.weight], dim=1)
    segs1 = [size_h] + (list(seg_widths) if blk.shortcut is not None else [])
    return {"w0": tc.pack(blk.fc_0.weight, seg_widths), "b0": blk.fc_0.bias.detach().float().contiguous(),
            "w1": tc.pack(w1, segs1), "b1": blk.fc_1.bias.detach().float().contiguous(),
            "size_h": size_h, "out": blk.fc_1.out_features, "shortcut": blk.shortcut is not None}


def _run_resblock(tc, pk, segs):
    """segs: [(tensor [M, ld], width[, group])] -- the block's input x as K segments (no ReLU flags: set here)."""
    h = tc.linear([(s[0], s[1], True) + tuple(s[2:]) for s in segs], pk["w0"], pk["size_h"], bias=pk["b0"])
    if pk["shortcut"]:
        return tc.linear([(h, pk["size_h"], True)] + [(s[0], s[1], False) + tuple(s[2:]) for s in segs], pk["w1"], pk["out"], bias=pk["b1"])
    if len(segs) != 1:
        raise RuntimeError("identity shortcut needs a single input segment")
    return tc.linear([(h, pk["size_h"], True)], pk["w1"], pk["out"], bias=pk["b1"], resid=segs[0][0])


# ------------------------------------------------------------------------------------------ ConvONet


class _Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.upconv = nn.ConvTranspose2d(cin, cout, 2, stride=2)
        self.conv1 = nn.Conv2d(2 * cout, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)

    def up(self, x):
        """ConvTranspose2d(k = 2, stride = 2) never overlaps its outputs: on the GPU it is one fp32 GEMM
        [B*H*W, Cin] x [Cin, Cout*4] followed by a pixel shuffle (cuDNN's fp32 path for it is a slow direct dgrad kernel:
        45 % of the encoder's time at B = 192)."""
        if not x.is_cuda:
            return self.upconv(x)
        B, cin, H, W = x.shape
        w = self.upconv.weight                                             # [Cin, Cout, 2, 2]
        cout = w.shape[1]
        y = x.permute(0, 2, 3, 1).reshape(B * H * W, cin) @ w.reshape(cin, cout * 4)
        y = y.view(B, H, W, cout, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, cout, 2 * H, 2 * W)
        return y + self.upconv.bias.view(1, cout, 1, 1)


class UNet(nn.Module):
    """2-D U-Net, concat merge, transpose-conv upsampling (the only mode the shipped config uses)."""

    def __init__(self, num_classes, in_channels=3, depth=5, start_filts=64, up_mode="transpose",
                 merge_mode="concat", **kwargs):
        super().__init__()
        if up_mode != "transpose" or merge_mode != "concat":
            raise RuntimeError("only up_mode='transpose', merge_mode='concat' is supported")
        self.depth = depth
        chans = [start_filts * 2 ** i for i in range(depth)]
        self.down_convs = nn.ModuleList(_Down(in_channels if i == 0 else chans[i - 1], chans[i]) for i in range(depth))
        self.up_convs = nn.ModuleList(_Up(chans[depth - 1 - i], chans[depth - 2 - i]) for i in range(depth - 1))
        self.conv_final = nn.Conv2d(chans[0], num_classes, 1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    def _pack(self, tc):
        def conv(c):                                                      # [Cout, Cin, 3, 3] -> [Cout, 9 Cin], k = (ky 3 + kx) Cin + ci
            return tc.pack(c.weight.permute(0, 2, 3, 1).reshape(c.out_channels, -1)), c.bias.detach().float().contiguous()

        def convt(c):                                                     # [Cin, Cout, 2, 2] -> [4 Cout, Cin], n = (i 2 + j) Cout + co
            return tc.pack(c.weight.permute(2, 3, 1, 0).reshape(4 * c.out_channels, c.in_channels)), c.bias.detach().float().contiguous()

        return {"down": [(conv(d.conv1), conv(d.conv2)) for d in self.down_convs],
                "up": [(convt(u.upconv), conv(u.conv1), conv(u.conv2)) for u in self.up_convs],
                "final": (tc.pack(self.conv_final.weight.reshape(self.conv_final.out_channels, -1)),
                          self.conv_final.bias.detach().float().contiguous())}

    def forward_cl(self, x, out=None):
        """The forward on channels-last tensors [B, H, W, C] with every convolution on the tcgen05 GEMM engine (csrc/tc_ops.cu):
        the 2x2 max-pool is taken on read by the next convolution, the skip concatenation is two K segments of the first
        up-convolution's GEMM, the transposed convolution is a GEMM with a pixel-shuffle epilogue.  -> [B, H, W, num_classes]
        (written into `out` when given).  unet.py:225-239."""
        from . import tc
        if not hasattr(self, "_tc_cache"):
            self._tc_cache = _PackedCache()
        pk = self._tc_cache.get(self, lambda: self._pack(tc))
        # every intermediate feature map lives in the engine's warp-transposed layout (tc.Blocked): only the engine touches them,
        # and its thread-per-pixel accesses are coalesced there; input and output stay plain channels-last
        skips = []
        for i, d in enumerate(self.down_convs):
            (w1, b1), (w2, b2) = pk["down"][i]
            x = tc.conv3x3(x, w1, b1, d.conv1.out_channels, pool=i > 0, blocked_out=True)
            x = tc.conv3x3(x, w2, b2, d.conv2.out_channels, blocked_out=True)
            skips.append(x)
        for i, u in enumerate(self.up_convs):
            (wt, bt), (w1, b1), (w2, b2) = pk["up"][i]
            B, H, W = x.dims
            cin, cout = x.C, u.upconv.out_channels
            up = tc.Blocked(B * 4 * H * W, cout, x.buf.device, dims=(B, 2 * H, 2 * W))
            tc.linear([(x, cin, False)], wt, 4 * cout, bias=bt, shuffle=(cout, H, W), out=up)
            x = tc.conv3x3(up, w1, b1, cout, src1=skips[-(i + 2)], blocked_out=True)
            x = tc.conv3x3(x, w2, b2, cout, blocked_out=True)
        B, H, W = x.dims
        wf, bf = pk["final"]
        ncls = self.conv_final.out_channels
        if out is None:
            out = torch.empty((B, H, W, ncls), dtype=torch.float32, device=x.buf.device)
        tc.linear([(x, x.C, False)], wf, ncls, bias=bf, out=out.view(B * H * W, ncls))
        return out

    def forward(self, x):
        skips = []
        for i, d in enumerate(self.down_convs):
            x = F.relu(d.conv2(F.relu(d.conv1(x))))
            skips.append(x)
            if i < self.depth - 1:
                x = F.max_pool2d(x, 2, 2)
        for i, u in enumerate(self.up_convs):
            x = torch.cat((u.up(x), skips[-(i + 2)]), 1)
            x = F.relu(u.conv2(F.relu(u.conv1(x))))
        return self.conv_final(x)


_PLANE_AXES = {"xz": [0, 2], "xy": [0, 1], "yz": [1, 2]}


def normalize_coordinate(p, padding=0.1, plane="xz"):
    """src/common.py:235-258, branch-free (no host sync): same values as the in-place masked writes."""
    xy = p[:, :, _PLANE_AXES[plane]] / (1 + padding + 10e-6) + 0.5
    xy = torch.where(xy >= 1, torch.full_like(xy, 1 - 10e-6), xy)
    return torch.where(xy < 0, torch.zeros_like(xy), xy)


def coordinate2index(x, reso):
    """src/common.py:300-315 ('2d')."""
    x = (x * reso).long()
    return (x[:, :, 0] + reso * x[:, :, 1])[:, None, :]


class LocalPoolPointnet(nn.Module):
    def __init__(self, c_dim=128, dim=3, hidden_dim=128, scatter_type="max", unet=False, unet_kwargs=None,
                 unet3d=False, unet3d_kwargs=None, plane_resolution=None, grid_resolution=None,
                 plane_type="xz", padding=0.1, n_blocks=5):
        super().__init__()
        if unet3d or "grid" in plane_type or scatter_type != "max":
            raise RuntimeError("only the shipped 3-plane / scatter_max configuration is supported")
        self.c_dim, self.hidden_dim = c_dim, hidden_dim
        self.fc_pos = nn.Linear(dim, 2 * hidden_dim)
        self.blocks = nn.ModuleList(ResnetBlockFC(2 * hidden_dim, hidden_dim) for _ in range(n_blocks))
        self.fc_c = nn.Linear(hidden_dim, c_dim)
        self.unet = UNet(c_dim, in_channels=c_dim, **unet_kwargs) if unet else None
        self.reso_plane, self.plane_type, self.padding = plane_resolution, list(plane_type), padding

    def _pool_local(self, index, c):
        """scatter_max into reso^2 bins then gather back, summed over planes (pointnet.py:104-122)."""
        src = c.permute(0, 2, 1)
        out = 0
        for pl in self.plane_type:
            idx = index[pl].expand(-1, src.shape[1], -1)
            binned = src.new_full((src.shape[0], src.shape[1], self.reso_plane ** 2), float("-inf"))
            binned.scatter_reduce_(-1, idx, src, reduce="amax", include_self=True)
            out = out + binned.gather(2, idx)
        return out.permute(0, 2, 1)

    def forward(self, p):
        """fp32 throughout (the reference runs fp32: TF32 convolutions -- torch's CUDA default -- would move the planes
        by ~1e-3 relative).  The scatter operators are the library's deterministic kernels and cuDNN's forward convolutions
        have no atomics, so the planes are reproducible run to run without torch's deterministic mode (which would route
        the U-Net through kernels 3-4x slower)."""
        if not p.is_cuda:
            return self._forward(p)
        tf32_c, tf32_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return self._forward(p)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_c, tf32_m

    def _forward(self, p):
        if p.is_cuda:
            return self._forward_cuda(p)
        index = {pl: coordinate2index(normalize_coordinate(p, self.padding, pl), self.reso_plane) for pl in self.plane_type}
        net = self.blocks[0](self.fc_pos(p))
        for block in self.blocks[1:]:
            net = block(torch.cat([net, self._pool_local(index, net)], dim=2))
        c = self.fc_c(net)
        fea = {}
        src = c.permute(0, 2, 1)
        for pl in self.plane_type:
            idx = index[pl].expand(-1, self.c_dim, -1)
            plane = src.new_zeros(p.size(0), self.c_dim, self.reso_plane ** 2).scatter_add_(-1, idx, src)
            cnt = torch.zeros_like(plane).scatter_add_(-1, idx, torch.ones_like(src))
            plane = (plane / cnt.clamp_(min=1)).reshape(p.size(0), self.c_dim, self.reso_plane, self.reso_plane)
            fea[pl] = self.unet(plane) if self.unet is not None else plane
        return fea

    def _forward_cuda(self, p):
        """The same forward on the GPU with every operator on the library's own kernels: bins, scatter_max + gather and
        scatter_mean (ifd_plane_bins, ifd_scatter_max_gather, ifd_scatter_mean_cl: deterministic), the Linear stack and the
        U-Net on the tcgen05 GEMM engine (tc.py; fp32 semantics through 3xTF32).  The three planes go through the shared U-Net
        as ONE batch of 3 B images and come out channels-last, the layout the decode kernel gathers from.  Nothing here
        depends on a cloud's position in the batch (no library picks an algorithm by batch size)."""
        from . import capi, tc
        capi.require_gpu()
        if list(self.plane_type) != ["xz", "xy", "yz"]:
            raise RuntimeError("the CUDA encoder path is built for plane_type ['xz', 'xy', 'yz']")
        L, st = capi.lib(), capi.stream()
        x = p.detach().float().contiguous()
        B, T, _ = x.shape
        R, nb, H = self.reso_plane, self.reso_plane ** 2, self.hidden_dim
        if not hasattr(self, "_tc_cache"):
            self._tc_cache = _PackedCache()

        def build():
            return {"pos": (tc.pack(self.fc_pos.weight), self.fc_pos.bias.detach().float().contiguous()),
                    "blocks": [_pack_resblock(tc, blk, [2 * H] if i == 0 else [H, H]) for i, blk in enumerate(self.blocks)],
                    "c": (tc.pack(self.fc_c.weight), self.fc_c.bias.detach().float().contiguous())}
        params = [q for m in (self.fc_pos, self.blocks, self.fc_c) for q in m.parameters()]
        pk = self._tc_cache.get(_ParamView(params), build)
        bins = torch.empty((3, B, T), dtype=torch.int32, device=x.device)
        capi.check(L.ifd_plane_bins(capi.ptr(x), B, T, R, float(self.padding), capi.ptr(bins), st), "ifd_plane_bins")
        M = B * T
        net0 = tc.linear([(x.view(M, 3), 3, False)], pk["pos"][0], 2 * H, bias=pk["pos"][1])
        net = _run_resblock(tc, pk["blocks"][0], [(net0, 2 * H)])
        for i in range(1, len(self.blocks)):
            pooled = torch.empty_like(net)
            capi.check(L.ifd_scatter_max_gather(capi.ptr(net), capi.ptr(bins), 3, B, T, H, nb, capi.ptr(pooled), st),
                       "ifd_scatter_max_gather")
            net = _run_resblock(tc, pk["blocks"][i], [(net, H), (pooled, H)])
        c = tc.linear([(net, H, False)], pk["c"][0], self.c_dim, bias=pk["c"][1])
        planes_in = torch.empty((3, B, R, R, self.c_dim), dtype=torch.float32, device=x.device)
        for i in range(3):
            capi.check(L.ifd_scatter_mean_cl(capi.ptr(c), capi.ptr(bins[i]), B, T, self.c_dim, nb, capi.ptr(planes_in[i]), st),
                       "ifd_scatter_mean_cl")
        if self.unet is not None:
            planes = self.unet.forward_cl(planes_in.view(3 * B, R, R, self.c_dim)).view(3, B, R, R, self.c_dim)
        else:
            planes = planes_in
        fea = PlaneFeatures((pl, planes[i].permute(0, 3, 1, 2)) for i, pl in enumerate(self.plane_type))   # [B,C,R,R] views
        fea.channels_last = planes                                        # the kernel layout [3,B,R,R,C], no copy needed
        return fea


class _ParamView:
    """Just enough of nn.Module for _PackedCache: a fixed list of parameters."""

    def __init__(self, params):
        self._params = list(params)

    def parameters(self):
        return self._params


class PlaneFeatures(dict):
    """encode_inputs' dict {'xz', 'xy', 'yz': [B, C, R, R]} whose three tensors are views of one channels-last buffer
    `channels_last` [3, B, R, R, C] -- what the decode kernels read (convonet.planes_to_channels_last takes it as is)."""
    channels_last = None


class LocalDecoder(nn.Module):
    """Parameter shell of the reference LocalDecoder; forward goes through the fused kernels."""

    def __init__(self, dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False, sample_mode="bilinear",
                 padding=0.1):
        super().__init__()
        if leaky or sample_mode != "bilinear":
            raise RuntimeError("only relu / bilinear is supported")
        self.c_dim, self.n_blocks, self.padding = c_dim, n_blocks, padding
        self.fc_c = nn.ModuleList(nn.Linear(c_dim, hidden_size) for _ in range(n_blocks))
        self.fc_p = nn.Linear(dim, hidden_size)
        self.blocks = nn.ModuleList(ResnetBlockFC(hidden_size) for _ in range(n_blocks))
        self.fc_out = nn.Linear(hidden_size, 1)
        self._packed = None

    def packed(self):
        """(Re)pack lazily so that load_state_dict after construction is honoured."""
        from . import convonet
        key = tuple(p._version for p in self.parameters()) + (str(next(self.parameters()).device),)
        if self._packed is None or self._packed[0] != key:
            sd = {"decoder." + k: v for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            self._packed = (key, convonet.ConvONetDecoder(sd, padding=self.padding, device=dev))
        return self._packed[1]

    def forward(self, p, c_plane, **kwargs):
        return self.packed().decode(p, c_plane).logits


class ConvolutionalOccupancyNetwork(nn.Module):
    def __init__(self, decoder, encoder=None, device=None):
        super().__init__()
        self.decoder = decoder.to(device)
        self.encoder = encoder.to(device) if encoder is not None else None
        self._device = device

    def encode_inputs(self, inputs):
        return self.encoder(inputs) if self.encoder is not None else torch.empty(inputs.size(0), 0)

    def decode(self, p, c, **kwargs):
        return dist.Bernoulli(logits=self.decoder(p, c, **kwargs))

    def forward(self, p, inputs, sample=True, **kwargs):
        return self.decode(p, self.encode_inputs(inputs), **kwargs)


# ------------------------------------------------------------------------------------------ ONet


class ResnetPointnet(nn.Module):
    def __init__(self, c_dim=128, dim=3, hidden_dim=128):
        super().__init__()
        self.fc_pos = nn.Linear(dim, 2 * hidden_dim)
        for i in range(5):
            setattr(self, "block_%d" % i, ResnetBlockFC(2 * hidden_dim, hidden_dim))
        self.fc_c = nn.Linear(hidden_dim, c_dim)

    def forward(self, p):
        if p.is_cuda:
            return self._forward_cuda(p)
        net = self.block_0(self.fc_pos(p))
        for i in range(1, 5):
            pooled = net.max(dim=1, keepdim=True)[0].expand(net.size())
            net = getattr(self, "block_%d" % i)(torch.cat([net, pooled], dim=2))
        return self.fc_c(F.relu(net.max(dim=1)[0]))

    def _forward_cuda(self, p):
        """ResnetPointnet.forward (ONet/im2mesh/encoder/pointnet.py:85-113) on the library's kernels: every Linear layer on
        the tcgen05 GEMM engine (tc.py, fp32 semantics), the global max-pool as ifd_group_max; the pooled vector is read as
        a per-cloud K segment instead of being expanded and concatenated."""
        from . import capi, tc
        capi.require_gpu()
        x = p.detach().float().contiguous()
        B, T, _ = x.shape
        blocks = [getattr(self, "block_%d" % i) for i in range(5)]
        H = blocks[0].fc_1.out_features
        if not hasattr(self, "_tc_cache"):
            self._tc_cache = _PackedCache()

        def build():
            return {"pos": (tc.pack(self.fc_pos.weight), self.fc_pos.bias.detach().float().contiguous()),
                    "blocks": [_pack_resblock(tc, blk, [2 * H] if i == 0 else [H, H]) for i, blk in enumerate(blocks)],
                    "c": (tc.pack(self.fc_c.weight), self.fc_c.bias.detach().float().contiguous())}
        pk = self._tc_cache.get(self, build)
        M = B * T
        net0 = tc.linear([(x.view(M, 3), 3, False)], pk["pos"][0], 2 * H, bias=pk["pos"][1])
        net = _run_resblock(tc, pk["blocks"][0], [(net0, 2 * H)])
        for i in range(1, 5):
            pooled = tc.group_max(net, B, T)                              # [B, H]
            net = _run_resblock(tc, pk["blocks"][i], [(net, H), (pooled, H, T)])
        pooled = tc.group_max(net, B, T)
        return tc.linear([(pooled, H, True)], pk["c"][0], self.fc_c.out_features, bias=pk["c"][1])


class CBatchNorm1d(nn.Module):
    def __init__(self, c_dim, f_dim, norm_method="batch_norm"):
        super().__init__()
        if norm_method != "batch_norm":
            raise RuntimeError("only batch_norm is supported")
        self.conv_gamma = nn.Conv1d(c_dim, f_dim, 1)
        self.conv_beta = nn.Conv1d(c_dim, f_dim, 1)
        self.bn = nn.BatchNorm1d(f_dim, affine=False)
        nn.init.zeros_(self.conv_gamma.weight)
        nn.init.zeros_(self.conv_beta.weight)
        nn.init.ones_(self.conv_gamma.bias)
        nn.init.zeros_(self.conv_beta.bias)

    def forward(self, x, c):
        c = c.unsqueeze(2) if c.dim() == 2 else c
        return self.conv_gamma(c) * self.bn(x) + self.conv_beta(c)


class CResnetBlockConv1d(nn.Module):
    def __init__(self, c_dim, size_in, size_h=None, size_out=None, norm_method="batch_norm", legacy=False):
        super().__init__()
        size_h = size_in if size_h is None else size_h
        size_out = size_in if size_out is None else size_out
        if legacy or size_in != size_out:
            raise RuntimeError("legacy CBN / shortcut blocks are not supported")
        self.bn_0 = CBatchNorm1d(c_dim, size_in, norm_method)
        self.bn_1 = CBatchNorm1d(c_dim, size_h, norm_method)
        self.fc_0 = nn.Conv1d(size_in, size_h, 1)
        self.fc_1 = nn.Conv1d(size_h, size_out, 1)
        nn.init.zeros_(self.fc_1.weight)

    def forward(self, x, c):
        net = self.fc_0(F.relu(self.bn_0(x, c)))
        return x + self.fc_1(F.relu(self.bn_1(net, c)))


class DecoderCBatchNorm(nn.Module):
    def __init__(self, dim=3, z_dim=128, c_dim=128, hidden_size=256, leaky=False, legacy=False):
        super().__init__()
        if z_dim != 0 or leaky or legacy:
            raise RuntimeError("only z_dim=0, relu, non-legacy is supported (configs/onet_mn40.yaml)")
        self.z_dim = z_dim
        self.fc_p = nn.Conv1d(dim, hidden_size, 1)
        for i in range(5):
            setattr(self, "block%d" % i, CResnetBlockConv1d(c_dim, hidden_size))
        self.bn = CBatchNorm1d(c_dim, hidden_size)
        self.fc_out = nn.Conv1d(hidden_size, 1, 1)

    def forward(self, p, z, c, **kwargs):
        """Plain-torch forward (used for the once-per-batch CBN folding checks; the loop uses the kernels)."""
        net = self.fc_p(p.transpose(1, 2))
        for i in range(5):
            net = getattr(self, "block%d" % i)(net, c)
        return self.fc_out(F.relu(self.bn(net, c))).squeeze(1)


class OccupancyNetwork(nn.Module):
    def __init__(self, decoder, encoder=None, encoder_latent=None, p0_z=None, device=None):
        super().__init__()
        self.decoder = decoder.to(device)
        self.encoder = encoder.to(device) if encoder is not None else None
        self._device = device

    def encode_inputs(self, inputs):
        return self.encoder(inputs) if self.encoder is not None else torch.empty(inputs.size(0), 0)

    def get_z_from_prior(self, size=torch.Size([]), sample=True):
        return torch.empty(*size, 0)                      # z_dim == 0 (onet_mn40.yaml:20)

    def decode(self, p, z, c, **kwargs):
        return dist.Bernoulli(logits=self.decoder(p, z, c, **kwargs))


# ------------------------------------------------------------------------------------------ builders

CONVONET_MN40 = dict(c_dim=32, padding=0.1, pointcloud_n=600, threshold=0.2,
                     encoder_kwargs=dict(hidden_dim=32, plane_type=["xz", "xy", "yz"], plane_resolution=64, unet=True,
                                         unet_kwargs=dict(depth=4, merge_mode="concat", start_filts=32)),
                     decoder_kwargs=dict(sample_mode="bilinear", hidden_size=32))
ONET_MN40 = dict(c_dim=512, pointcloud_n=300, threshold=0.2, encoder_kwargs=dict(hidden_dim=512), z_dim=0)


def build_convonet(cfg=None, device=None):
    """configs/convonet_3plane_mn40.yaml -> model (src/conv_onet/config.py:8-84)."""
    cfg = cfg or CONVONET_MN40
    dec = LocalDecoder(dim=3, c_dim=cfg["c_dim"], padding=cfg["padding"], **cfg["decoder_kwargs"])
    enc = LocalPoolPointnet(dim=3, c_dim=cfg["c_dim"], padding=cfg["padding"], **cfg["encoder_kwargs"])
    return ConvolutionalOccupancyNetwork(dec, enc, device=device)


def build_onet(cfg=None, device=None):
    """configs/onet_mn40.yaml -> model (im2mesh/onet/config.py)."""
    cfg = cfg or ONET_MN40
    dec = DecoderCBatchNorm(dim=3, z_dim=cfg["z_dim"], c_dim=cfg["c_dim"])
    enc = ResnetPointnet(dim=3, c_dim=cfg["c_dim"], **cfg["encoder_kwargs"])
    return OccupancyNetwork(dec, enc, device=device)


def randomize_zero_init_(sd, seed=0):
    """Synthetic-weights protocol: give every tensor the reference zero-initialises a
    non-trivial value so that all terms of the path are exercised.  In place; returns sd."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k.endswith("fc_1.weight"):
            fan_in = v.shape[1] * (v.shape[2] if v.dim() > 2 else 1)
            v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) / math.sqrt(fan_in))
        elif k.endswith("conv_gamma.weight") or k.endswith("conv_beta.weight"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.02)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return sd


def synthetic_state_dict(kind="convonet", seed=0):
    """Deterministic random-init weights with the reference's names and shapes (no checkpoint ships with the
    reference: ConvONet/README.md:17, ONet/README.md:17)."""
    torch.manual_seed(seed)
    model = build_convonet() if kind == "convonet" else build_onet()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    return randomize_zero_init_(sd, seed)
