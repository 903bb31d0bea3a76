"""Dense layers of the encoders on the tensor cores: thin wrappers over ifd_tc_* (include/ifd_b200.h), fp32 semantics
(3xTF32).  Tensors stay torch tensors; only data_ptr() and the current stream cross the C ABI."""
import ctypes

import torch

from . import capi


class Blocked:
    """A [rows, C] tensor in the engine's warp-transposed layout (include/ifd_b200.h, ifd_tc_blocked_floats): rows in blocks
    of 32, inside a block the C / 4 float4 column groups one after the other.  Only the engine reads it; `dims` carries the
    logical image shape (B, H, W) when the rows are pixels."""

    def __init__(self, rows, C, device, dims=None):
        self.rows, self.C, self.dims = int(rows), int(C), dims
        self.buf = torch.empty(int(capi.lib().ifd_tc_blocked_floats(self.rows, self.C)), dtype=torch.float32, device=device)

    @staticmethod
    def from_rows(t, dims=None):
        """plain [rows, C] (or channels-last [B, H, W, C]) tensor -> Blocked copy (tests / debugging)."""
        if t.dim() == 4:
            dims = tuple(t.shape[:3])
            t = t.reshape(-1, t.shape[3])
        rows, C = t.shape
        b = Blocked(rows, C, t.device, dims=dims)
        nb = (rows + 31) // 32
        pad = torch.zeros((nb * 32, C), dtype=torch.float32, device=t.device)
        pad[:rows] = t
        b.buf.copy_(pad.view(nb, 32, C // 4, 4).permute(0, 2, 1, 3).reshape(-1))
        return b

    def to_rows(self):
        """-> plain [rows, C] tensor (tests / debugging)."""
        nb = (self.rows + 31) // 32
        return self.buf.view(nb, self.C // 4, 32, 4).permute(0, 2, 1, 3).reshape(nb * 32, self.C)[:self.rows]


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
        if isinstance(t, Blocked):
            if width != t.C or group:
                raise RuntimeError("a blocked segment is used whole and ungrouped")
            M = t.rows if M is None else M
            if t.rows != M:
                raise RuntimeError("A segments disagree on the number of rows")
            a.a_ptr[i], a.a_ld[i], a.a_width[i], a.a_relu[i], a.a_group[i], a.a_blocked[i] = t.buf.data_ptr(), t.C, t.C, int(bool(relu)), 0, 1
            continue
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
    first = segments[0][0]
    dev = first.buf.device if isinstance(first, Blocked) else first.device
    if shuffle is not None:
        cout, H, W = shuffle
        if out is None:
            out = torch.empty((M // (H * W), 2 * H, 2 * W, cout), dtype=torch.float32, device=dev)
        a.shuffle_cout, a.shuffle_H, a.shuffle_W = cout, H, W
        if isinstance(out, Blocked):
            a.out, a.ld_out, a.out_blocked = out.buf.data_ptr(), cout, 1
        else:
            a.out, a.ld_out = out.data_ptr(), cout
    elif isinstance(out, Blocked):
        a.out, a.ld_out, a.out_blocked = out.buf.data_ptr(), N, 1
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=dev)
        a.out, a.ld_out = out.data_ptr(), out.stride(0)
    capi.check(capi.lib().ifd_tc_linear(ctypes.byref(a), capi.stream()), "ifd_tc_linear")
    return out


def conv3x3(src0, wimg, bias, cout, src1=None, pool=False, relu=True, out=None, blocked_out=False):
    """Inputs: channels-last [B, H, W, C] tensors or `Blocked` feature maps (dims = (B, H, W)); pool: src0 is the [B, 2H, 2W, C]
    map and is max-pooled 2x2 on read.  -> out: a channels-last tensor, or a Blocked map when blocked_out (or `out` is one)."""
    def dims(t):
        return (t.dims + (t.C,)) if isinstance(t, Blocked) else tuple(t.shape)
    B, H, W, C0 = dims(src0)
    if pool:
        H, W = H // 2, W // 2
    C1 = dims(src1)[3] if src1 is not None else 0
    dev = src0.buf.device if isinstance(src0, Blocked) else src0.device
    if out is None:
        out = Blocked(B * H * W, cout, dev, dims=(B, H, W)) if blocked_out else torch.empty((B, H, W, cout), dtype=torch.float32, device=dev)
    layout = (1 if isinstance(src0, Blocked) else 0) | (2 if isinstance(src1, Blocked) else 0) | (4 if isinstance(out, Blocked) else 0)
    ptr = lambda t: None if t is None else (t.buf.data_ptr() if isinstance(t, Blocked) else capi.ptr(t))
    capi.check(capi.lib().ifd_tc_conv3x3(ptr(src0), C0, ptr(src1), C1, B, H, W, int(bool(pool)), capi.ptr(wimg), capi.ptr(bias),
                                         int(bool(relu)), cout, ptr(out), layout, capi.stream()), "ifd_tc_conv3x3")
    return out


def group_max(x, groups, T):
    """x [groups * T, C] -> [groups, C]."""
    C = x.shape[1]
    out = torch.empty((groups, C), dtype=torch.float32, device=x.device)
    capi.check(capi.lib().ifd_group_max(capi.ptr(x), groups, T, C, capi.ptr(out), capi.stream()), "ifd_group_max")
    return out
