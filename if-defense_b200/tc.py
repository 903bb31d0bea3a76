"""Dense layers of the encoders on the tensor cores: thin wrappers over ifd_tc_* (include/ifd_b200.h), fp32 semantics
(3xTF32).  Tensors stay torch tensors; only data_ptr() and the current stream cross the C ABI."""
import ctypes

import torch

from . import capi


def pack(weight2d, seg_widths=None):
    """nn.Linear-style weight [N, K] (cuda, float32) -> packed image tensor for ifd_tc_linear / ifd_tc_conv3x3.  seg_widths: the
    K split of a concatenated input (defaults to one segment)."""
    w = weight2d.detach().float().contiguous()
    N, K = w.shape
    seg = list(seg_widths) if seg_widths is not None else [K]
    if sum(seg) != K:
        raise RuntimeError("segment widths do not add up to K")
    arr = (ctypes.c_int * len(seg))(*seg)
    L = capi.lib()
    n = L.ifd_tc_packed_floats(N, arr, len(seg))
    out = torch.empty(n, dtype=torch.float32, device=w.device)
    capi.check(L.ifd_tc_pack(capi.ptr(w), N, arr, len(seg), K, capi.ptr(out), capi.stream()), "ifd_tc_pack")
    return out


def linear(segments, wimg, N, bias=None, resid=None, relu_out=False, out=None, shuffle=None):
    """segments: list of (tensor [M, ld] or [G, ld] cuda float32 row-major, width, relu_on_read[, group]).
    -> out [M, N] (or the given `out`, a 2-D view whose row stride is its ld).  shuffle = (cout, H, W): ConvTranspose2d(2, 2)
    epilogue, `out` must then be the [B, 2H, 2W, cout] channels-last tensor."""
    a = capi.TcLinearArgs()
    M = None
    for i, seg in enumerate(segments):
        t, width, relu = seg[0], seg[1], seg[2]
        group = seg[3] if len(seg) > 3 else 0
        if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float32:
            raise RuntimeError("A segments must be 2-D float32 row-major tensors")
        rows = t.shape[0] * (group if group else 1)
        M = rows if M is None else M
        if rows != M:
            raise RuntimeError("A segments disagree on the number of rows")
        a.a_ptr[i], a.a_ld[i], a.a_width[i], a.a_relu[i], a.a_group[i] = t.data_ptr(), t.stride(0), int(width), int(bool(relu)), int(group)
    a.M, a.N, a.n_seg = M, N, len(segments)
    a.wimg = wimg.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    if resid is not None:
        a.resid, a.ld_resid = resid.data_ptr(), resid.stride(0)
    a.relu_out = int(bool(relu_out))
    dev = segments[0][0].device
    if shuffle is not None:
        cout, H, W = shuffle
        if out is None:
            out = torch.empty((M // (H * W), 2 * H, 2 * W, cout), dtype=torch.float32, device=dev)
        a.shuffle_cout, a.shuffle_H, a.shuffle_W = cout, H, W
        a.out, a.ld_out = out.data_ptr(), cout
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=dev)
        a.out, a.ld_out = out.data_ptr(), out.stride(0)
    capi.check(capi.lib().ifd_tc_linear(ctypes.byref(a), capi.stream()), "ifd_tc_linear")
    return out


def conv3x3(src0, wimg, bias, cout, src1=None, pool=False, relu=True, out=None):
    """channels-last [B, H, W, C] tensors; pool: src0 is [B, 2H, 2W, C] and is max-pooled 2x2 on read."""
    B, H, W, C0 = src0.shape
    if pool:
        H, W = H // 2, W // 2
    C1 = src1.shape[3] if src1 is not None else 0
    if out is None:
        out = torch.empty((B, H, W, cout), dtype=torch.float32, device=src0.device)
    capi.check(capi.lib().ifd_tc_conv3x3(capi.ptr(src0), C0, capi.ptr(src1), C1, B, H, W, int(bool(pool)), capi.ptr(wimg), capi.ptr(bias),
                                         int(bool(relu)), cout, capi.ptr(out), capi.stream()), "ifd_tc_conv3x3")
    return out


def group_max(x, groups, T):
    """x [groups * T, C] -> [groups, C]."""
    C = x.shape[1]
    out = torch.empty((groups, C), dtype=torch.float32, device=x.device)
    capi.check(capi.lib().ifd_group_max(capi.ptr(x), groups, T, C, capi.ptr(out), capi.stream()), "ifd_group_max")
    return out
