"""ConvONet-Opt on the sm_100a kernels, behind the reference's seams.

  ConvONetDecoder.decode(p, c).logits     <- generator.model.decode(p, c).logits   (ConvONet/opt_defense.py:212,
                                             src/conv_onet/models/__init__.py:67-77)
  Restorer.optimize_points(...)           <- optimize_points(opt_points, z, c, rep_weight, iterations, printing)
                                             (ConvONet/opt_defense.py:182-239)
  normalize_batch_pc / init_points        <- opt_defense.py:76-83 / :149-179

`c` is exactly what the reference's encode_inputs returns: a dict {'xz','xy','yz': [B,32,64,64]}.
"""
import ctypes

import numpy as np
import torch
from torch import distributions as dist

from . import capi, weights

PLANES = ("xz", "xy", "yz")


def planes_to_channels_last(c):
    """dict of [B,C,R,R] (encoder output) -> one device tensor [3,B,R,R,C] in kernel layout."""
    if isinstance(c, torch.Tensor):          # already converted
        return c
    cl = getattr(c, "channels_last", None)
    if cl is not None:                       # models.LocalPoolPointnet on CUDA: the planes are born in the kernel layout
        return cl
    capi.require_gpu()
    if all(c[k].dim() == 4 and c[k].dtype == torch.float32 and c[k].is_contiguous(memory_format=torch.channels_last)
           and not c[k].is_contiguous() for k in PLANES):
        # the encoder already produced channels_last memory (models.LocalPoolPointnet on CUDA): that IS the kernel layout
        return torch.stack([c[k].detach().permute(0, 2, 3, 1) for k in PLANES]).contiguous()
    x = torch.stack([c[k].detach().float() for k in PLANES]).contiguous()     # [3,B,C,R,R]
    _, B, C, R, R2 = x.shape
    if R != R2:
        raise RuntimeError("feature planes must be square")
    out = torch.empty((3, B, R, R, C), dtype=torch.float32, device=x.device)
    capi.check(capi.lib().ifd_planes_nchw_to_cl(capi.ptr(x), capi.ptr(out), 3 * B, C, R, capi.stream()),
               "ifd_planes_nchw_to_cl")
    return out


def volume_to_channels_last(vol):
    """'grid' feature volume [B,C,R,R,R] (what the reference's generate_grid_features returns, indexed [z][y][x]) -> the
    kernel layout [B,R,R,R,C]; a 5-D channels_last_3d tensor is taken as it is."""
    if vol.dim() != 5:
        raise RuntimeError("the grid feature volume must be [B,C,R,R,R]")
    return vol.detach().float().permute(0, 2, 3, 4, 1).contiguous()


class _GridDecodeFn(torch.autograd.Function):
    """generator.model.decode(p, {'grid': volume}).logits, differentiable w.r.t. p (decoder.py:59-67,72-73)."""

    @staticmethod
    def forward(ctx, p, vol_cl, wblob, dims, padding):
        capi.require_gpu()
        x = p.detach().float().contiguous()
        B, K, _ = x.shape
        C, H, nb = dims
        logits = torch.empty((B, K), dtype=torch.float32, device=x.device)
        capi.check(capi.lib().ifd_convonet_grid_decode_fwd(capi.ptr(vol_cl), capi.ptr(wblob), capi.ptr(x), B, K, vol_cl.shape[1], C, H,
                                                           nb, padding, capi.ptr(logits), capi.stream()), "ifd_convonet_grid_decode_fwd")
        ctx.save_for_backward(x, vol_cl, wblob)
        ctx.dims, ctx.padding = dims, padding
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        x, vol_cl, wblob = ctx.saved_tensors
        B, K, _ = x.shape
        C, H, nb = ctx.dims
        g = grad_logits.detach().float().contiguous()
        out = torch.empty_like(x)
        capi.check(capi.lib().ifd_convonet_grid_decode_bwd(capi.ptr(vol_cl), capi.ptr(wblob), capi.ptr(x), capi.ptr(g), B, K,
                                                           vol_cl.shape[1], C, H, nb, ctx.padding, capi.ptr(out), capi.stream()),
                   "ifd_convonet_grid_decode_bwd")
        return out, None, None, None, None


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, planes_cl, wblob, dims, padding):
        capi.require_gpu()
        x = p.detach().float().contiguous()
        B, K, _ = x.shape
        C, H, nb = dims
        R = planes_cl.shape[2]
        logits = torch.empty((B, K), dtype=torch.float32, device=x.device)
        capi.check(capi.lib().ifd_convonet_decode_fwd(capi.ptr(planes_cl), capi.ptr(wblob), capi.ptr(x), B, K, R, C, H, nb,
                                                      padding, capi.ptr(logits), capi.stream()), "ifd_convonet_decode_fwd")
        ctx.save_for_backward(x, planes_cl, wblob)
        ctx.dims, ctx.padding = dims, padding
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        x, planes_cl, wblob = ctx.saved_tensors
        B, K, _ = x.shape
        C, H, nb = ctx.dims
        R = planes_cl.shape[2]
        g = grad_logits.detach().float().contiguous()
        out = torch.empty_like(x)
        capi.check(capi.lib().ifd_convonet_decode_bwd(capi.ptr(planes_cl), capi.ptr(wblob), capi.ptr(x), capi.ptr(g), B, K, R,
                                                      C, H, nb, ctx.padding, capi.ptr(out), capi.stream()),
                   "ifd_convonet_decode_bwd")
        return out, None, None, None, None


class ConvONetDecoder:
    """The decoder half of ConvolutionalOccupancyNetwork, frozen, with weights packed for the kernels.

    Built from the reference checkpoint interface: any state_dict with the `decoder.*` keys of
    ConvONet/src/conv_onet/models/decoder.py."""

    def __init__(self, state_dict, padding=0.1, device="cuda", prefix="decoder."):
        self.dims = weights.convonet_decoder_dims(state_dict, prefix)
        self.padding = float(padding)
        blob = weights.pack_convonet_decoder(state_dict, prefix)
        self.blob_host = blob
        self.device = capi.use_device(device)
        self.blob = torch.from_numpy(blob).to(self.device) if self.device.type == "cuda" else None

    def decode(self, p, c, **kwargs):
        if isinstance(c, dict) and "grid" in c:
            if len(c) != 1:
                raise RuntimeError("a feature volume combined with planes is not supported (no reference config does it)")
            logits = _GridDecodeFn.apply(p, volume_to_channels_last(c["grid"]), self.blob, self.dims, self.padding)
            return dist.Bernoulli(logits=logits)
        planes = planes_to_channels_last(c)
        logits = _DecodeFn.apply(p, planes, self.blob, self.dims, self.padding)
        return dist.Bernoulli(logits=logits)


def normalize_batch_pc(points):
    """opt_defense.py:76-83 (torch ops; the fused loop does this on device itself)."""
    points -= torch.mean(points, dim=1)[:, None, :]
    d = torch.sum(points ** 2, dim=2) ** 0.5
    points /= torch.max(d, dim=1)[0][:, None, None]
    return points


class Restorer:
    """Holds what the reference keeps in module globals (generator, args.threshold, args.lr) and exposes
    optimize_points with the reference's signature."""

    def __init__(self, decoder, threshold=0.2, lr=1e-3, decode_kernel=0, side_by_side=True):
        """decode_kernel: 0 = production default (tcgen05 tensor-core MLP, 3xTF32); 2 = fp32 SIMT kernels, for
        bit-level trajectory comparisons with an fp32 reference (see include/ifd_b200.h, ifd_opt_params).
        side_by_side: a batch of >= 96 clouds is cut into equal parts of at most 64 clouds whose loops run next to each other
        (ifd_convonet_opt_batches with the batch's B_ref: the same bits, the launches of one loop fill the SMs the others
        leave idle); an int forces the number of parts."""
        self.decoder, self.threshold, self.lr = decoder, float(threshold), float(lr)
        self.decode_kernel = int(decode_kernel)
        self.side_by_side = side_by_side
        self.last_stats = None

    def _parts(self, B):
        if not self.side_by_side or B < 96:
            return 1
        if isinstance(self.side_by_side, int) and not isinstance(self.side_by_side, bool):
            return self.side_by_side if B % self.side_by_side == 0 else 1
        for n in range(2, 9):                 # parts of at most 64 clouds: 192 -> 3 x 64, 164 -> 4 x 41, 128 -> 2 x 64
            if B % n == 0 and B // n <= 64:
                return n
        return 1

    def params(self, B_ref, rep_weight, iterations, want_stats=False, **over):
        return capi.default_params(n_steps=iterations + 1, B_ref=int(B_ref), rep_weight=float(rep_weight),
                                   lr=self.lr, occ_target=self.threshold, padding=self.decoder.padding,
                                   want_stats=int(bool(want_stats)), decode_kernel=self.decode_kernel, **over)

    def optimize_points(self, opt_points, z, c, rep_weight=1., iterations=1000, printing=False, B_ref=None,
                        return_tensor=False):
        """Same arguments as the reference.  `B_ref` (extension): the batch size of the reference call these
        clouds belong to, when a batch is split across calls / GPUs (defaults to len(opt_points))."""
        capi.require_gpu()
        x = opt_points.detach().float().cuda().contiguous().clone()
        B, K, _ = x.shape
        if isinstance(c, dict) and "grid" in c:
            return self._optimize_points_grid(x, c, rep_weight, iterations, printing, B if B_ref is None else B_ref, return_tensor)
        planes = planes_to_channels_last(c)
        C, H, nb = self.decoder.dims
        R = planes.shape[2]
        L = capi.lib()
        P = self.params(B if B_ref is None else B_ref, rep_weight, iterations, want_stats=printing)
        parts = 1 if printing else self._parts(B)
        if parts > 1:
            Bc = B // parts
            ws = torch.empty(L.ifd_convonet_opt_batches_workspace_bytes(Bc, K), dtype=torch.uint8, device=x.device)
            pls = [planes[:, i * Bc:(i + 1) * Bc].contiguous() for i in range(parts)]
            pp = (ctypes.c_void_p * parts)(*[t.data_ptr() for t in pls])
            xp = (ctypes.c_void_p * parts)(*[x[i * Bc:(i + 1) * Bc].data_ptr() for i in range(parts)])
            capi.check(L.ifd_convonet_opt_batches(parts, pp, capi.ptr(self.decoder.blob), xp, Bc, K, R, C, H, nb, ctypes.byref(P),
                                                  capi.ptr(ws), ws.numel(), capi.stream()), "ifd_convonet_opt_batches")
            return x if return_tensor else x.cpu().numpy()
        ws_bytes = L.ifd_convonet_opt_workspace_bytes(B, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        n_stat = iterations // 100 + 1
        stats = torch.zeros((n_stat, 4), dtype=torch.float64, device=x.device) if printing else None
        capi.check(L.ifd_convonet_opt(capi.ptr(planes), capi.ptr(self.decoder.blob), capi.ptr(x), None, None, B, K, R, C, H,
                                      nb, ctypes.byref(P), capi.ptr(stats), capi.ptr(ws), ws_bytes, capi.stream()),
                   "ifd_convonet_opt")
        if printing:
            self.last_stats = stats.cpu().numpy()
            for j, (loss, occ, rep, sg) in enumerate(self.last_stats):
                print('iter {}, loss {:.4f}'.format(j * 100, loss))
                print('occ loss: {:.4f}, rep loss: {:.4f}\nocc value mean: {:.4f}'.format(occ, rep, sg))
        if return_tensor:
            return x
        return x.cpu().numpy()

    def _optimize_points_grid(self, x, c, rep_weight, iterations, printing, B_ref, return_tensor):
        """The loop with the 'grid' decoder (c = {'grid': [B,C,R,R,R]}): ifd_convonet_grid_opt."""
        if len(c) != 1:
            raise RuntimeError("a feature volume combined with planes is not supported (no reference config does it)")
        vol = volume_to_channels_last(c["grid"])
        B, K, _ = x.shape
        C, H, nb = self.decoder.dims
        L = capi.lib()
        P = self.params(B_ref, rep_weight, iterations, want_stats=printing)
        ws_bytes = L.ifd_convonet_opt_workspace_bytes(B, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        stats = torch.zeros((iterations // 100 + 1, 4), dtype=torch.float64, device=x.device) if printing else None
        capi.check(L.ifd_convonet_grid_opt(capi.ptr(vol), capi.ptr(self.decoder.blob), capi.ptr(x), None, None, B, K, vol.shape[1], C, H,
                                           nb, ctypes.byref(P), capi.ptr(stats), capi.ptr(ws), ws_bytes, capi.stream()),
                   "ifd_convonet_grid_opt")
        if printing:
            self.last_stats = stats.cpu().numpy()
        return x if return_tensor else x.cpu().numpy()

    def optimize_points_host(self, opt_points_np, planes_nchw_np, rep_weight=500., iterations=200, B_ref=None):
        """End-to-end seam with HOST buffers (numpy in, numpy out): H2D, layout conversion, loop, D2H inside
        one C call (ifd_convonet_opt_host).  planes_nchw_np: [3,B,C,R,R] float32."""
        capi.require_gpu()
        x = np.ascontiguousarray(opt_points_np, dtype=np.float32).copy()
        pl = np.ascontiguousarray(planes_nchw_np, dtype=np.float32)
        B, K, _ = x.shape
        _, _, C, R, _ = pl.shape
        Cd, H, nb = self.decoder.dims
        P = self.params(B if B_ref is None else B_ref, rep_weight, iterations)
        capi.check(capi.lib().ifd_convonet_opt_host(pl.ctypes.data, self.decoder.blob_host.ctypes.data, x.ctypes.data, B, K, R,
                                                    C, H, nb, ctypes.byref(P), None), "ifd_convonet_opt_host")
        return x

    def optimize_points_host_many(self, opt_points_list, planes_nchw_list, rep_weight=500., iterations=200, B_ref=None,
                                  out=None):
        """The batch loop of defend_point_cloud (opt_defense.py:272-312) through one pipelined C call
        (ifd_convonet_opt_host_batches): lists of HOST arrays in ([B,K,3] init points, [3,B,C,R,R] planes per batch,
        all batches the same shape; pinned memory -- e.g. torch.Tensor.pin_memory().numpy() -- makes the copies
        overlap the previous batch's loop), list of restored [B,K,3] arrays out.  `out` (optional): a list of float32
        [B,K,3] arrays in pinned memory to restore into (saves the pinned allocations of this call)."""
        capi.require_gpu()
        if len(opt_points_list) != len(planes_nchw_list):
            raise RuntimeError("need one planes array per batch")
        if not opt_points_list:
            return []
        # in/out point buffers in pinned memory (torch's caching host allocator): a D2H into pageable memory would
        # block the host inside the batch loop and serialise the pipeline
        if out is None:
            keep = [torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)).clone().pin_memory() for p in opt_points_list]
            outs = [t.numpy() for t in keep]
        else:
            if len(out) != len(opt_points_list):
                raise RuntimeError("need one output array per batch")
            outs = list(out)
            for o, p in zip(outs, opt_points_list):
                if o.dtype != np.float32 or not o.flags["C_CONTIGUOUS"] or o.shape != np.shape(p):
                    raise RuntimeError("output arrays must be C-contiguous float32 of the input shape")
                np.copyto(o, p)
        pls = [np.ascontiguousarray(p, dtype=np.float32) for p in planes_nchw_list]
        B, K, _ = outs[0].shape
        _, _, C, R, _ = pls[0].shape
        if any(o.shape != (B, K, 3) for o in outs) or any(p.shape != (3, B, C, R, R) for p in pls):
            raise RuntimeError("all batches of one call must have the same shape")
        _, H, nb = self.decoder.dims
        P = self.params(B if B_ref is None else B_ref, rep_weight, iterations)
        n = len(outs)
        pp = (ctypes.c_void_p * n)(*[p.ctypes.data for p in pls])
        xp = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
        capi.check(capi.lib().ifd_convonet_opt_host_batches(n, pp, self.decoder.blob_host.ctypes.data, xp, B, K, R, C, H, nb,
                                                            ctypes.byref(P)), "ifd_convonet_opt_host_batches")
        return outs
