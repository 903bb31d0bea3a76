"""ONet-Opt on the sm_100a kernels, behind the reference's seams.

  ONetDecoder.decode(p, z, c).logits      <- generator.model.decode(opt_points, z, c).logits   (ONet/opt_defense.py:212)
  ONetRestorer.optimize_points(...)       <- optimize_points(opt_points, z, c, rep_weight, iterations, printing)
                                             (ONet/opt_defense.py:182-239)

`c` is what the reference's encode_inputs returns for ONet: a [B,512] tensor; `z` is the empty [B,0] prior sample
(z_dim = 0) and is ignored.
"""
import ctypes

import torch
from torch import distributions as dist

from . import capi, weights


class _Workspace:
    """Caches the (large) ONet workspace per (B, K): 11 saved activation tensors + weight images."""

    def __init__(self):
        self.key, self.buf = None, None

    def get(self, B, K, device):
        if self.key != (B, K, str(device)):
            n = capi.lib().ifd_onet_workspace_bytes(B, K)
            self.buf = torch.empty(n, dtype=torch.uint8, device=device)
            self.key = (B, K, str(device))
        return self.buf


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, c, dec):
        capi.require_gpu()
        x = p.detach().float().contiguous()
        cc = c.detach().float().contiguous()
        B, K, _ = x.shape
        L = capi.lib()
        ws = dec.ws.get(B, K, x.device)
        capi.check(L.ifd_onet_prepare(capi.ptr(dec.blob), capi.ptr(cc), B, K, capi.ptr(ws), ws.numel(), capi.stream()), "ifd_onet_prepare")
        logits = torch.empty((B, K), dtype=torch.float32, device=x.device)
        capi.check(L.ifd_onet_decode_fwd(capi.ptr(dec.blob), capi.ptr(x), B, K, capi.ptr(logits), capi.ptr(ws), ws.numel(),
                                         capi.stream()), "ifd_onet_decode_fwd")
        ctx.save_for_backward(x, cc)
        ctx.dec = dec
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        x, cc = ctx.saved_tensors
        dec = ctx.dec
        B, K, _ = x.shape
        L = capi.lib()
        ws = dec.ws.get(B, K, x.device)
        capi.check(L.ifd_onet_prepare(capi.ptr(dec.blob), capi.ptr(cc), B, K, capi.ptr(ws), ws.numel(), capi.stream()), "ifd_onet_prepare")
        g = grad_logits.detach().float().contiguous()
        out = torch.empty_like(x)
        capi.check(L.ifd_onet_decode_bwd(capi.ptr(dec.blob), capi.ptr(x), capi.ptr(g), B, K, capi.ptr(out), capi.ptr(ws), ws.numel(),
                                         capi.stream()), "ifd_onet_decode_bwd")
        return out, None, None


class ONetDecoder:
    """DecoderCBatchNorm, frozen, eval mode, packed for the kernels (checkpoint interface: the reference's
    `decoder.*` state_dict keys)."""

    def __init__(self, state_dict, device="cuda", prefix="decoder."):
        self.blob_host = weights.pack_onet_decoder(state_dict, prefix)
        self.device = capi.use_device(device)
        self.blob = torch.from_numpy(self.blob_host).to(self.device) if self.device.type == "cuda" else None
        self.ws = _Workspace()

    def decode(self, p, z, c, **kwargs):
        return dist.Bernoulli(logits=_DecodeFn.apply(p, c, self))

    def _prepare_eval(self, cc, k_max):
        """ifd_onet_prepare for ONE shape (cc [1,512] cuda) on a workspace that holds chunks of up to k_max points.  The
        prepared state (weight images, folded CBN scale/shift) sits at the head of the workspace and depends on B only, so
        every later ifd_onet_decode_fwd with K <= k_max can reuse it."""
        ws = self.ws.get(1, k_max, cc.device)
        capi.check(capi.lib().ifd_onet_prepare(capi.ptr(self.blob), capi.ptr(cc), 1, k_max, capi.ptr(ws), ws.numel(), capi.stream()),
                   "ifd_onet_prepare")
        return ws

    def _logits_prepared(self, x, ws):
        """x [1,K,3] cuda -> logits [K] cuda (forward only, nothing saved); ws from _prepare_eval with k_max >= K."""
        K = x.shape[1]
        logits = torch.empty((1, K), dtype=torch.float32, device=x.device)
        capi.check(capi.lib().ifd_onet_decode_fwd(capi.ptr(self.blob), capi.ptr(x), 1, K, capi.ptr(logits), capi.ptr(ws), ws.numel(),
                                                  capi.stream()), "ifd_onet_decode_fwd")
        return logits[0]

    def _logits_nograd(self, x, cc):
        """x [1,K,3] cuda, cc [1,512] cuda -> logits [K] cuda."""
        return self._logits_prepared(x, self._prepare_eval(cc, x.shape[1]))

    def eval_points(self, p, z=None, c=None, points_batch_size=100000):
        """Generator3D.eval_points (ONet/im2mesh/onet/generation.py:138-158): occupancy logits of N points for ONE shape,
        in chunks of points_batch_size (the reference's default, configs/default.yaml), returned on the CPU.  The weight
        images and the CBN fold are prepared once per call, not per chunk."""
        capi.require_gpu()
        p = torch.as_tensor(p).detach().float()
        cc = torch.as_tensor(c).detach().float().reshape(1, -1).cuda().contiguous()
        out = torch.empty(p.shape[0], dtype=torch.float32)
        if p.shape[0] == 0:
            return out
        ws = self._prepare_eval(cc, min(points_batch_size, p.shape[0]))
        for lo in range(0, p.shape[0], points_batch_size):
            x = p[lo:lo + points_batch_size].cuda().contiguous().view(1, -1, 3)
            out[lo:lo + x.shape[1]] = self._logits_prepared(x, ws).cpu()
        return out

    def eval_lattice_points(self, points, resolution, cc, ws, padding=0.1, points_batch_size=100000):
        """Logits (float32 cuda [N]) of integer lattice points [N,3] (cuda int64) of the `resolution` lattice, in the
        reference's coordinates: box_size * (points / resolution - 0.5) in float32 (generation.py:121-124)."""
        pointsf = (1.0 + padding) * (points.float() / resolution - 0.5)
        values = torch.empty(points.shape[0], dtype=torch.float32, device=points.device)
        for lo in range(0, points.shape[0], points_batch_size):
            x = pointsf[lo:lo + points_batch_size].contiguous().view(1, -1, 3)
            values[lo:lo + x.shape[1]] = self._logits_prepared(x, ws)
        return values

    def eval_dense_grid(self, c, resolution=128, padding=0.1, points_batch_size=100000):
        """The occupancy field generate_from_latent needs (ONet/im2mesh/onet/generation.py:101-130), evaluated densely
        on the (resolution + 1)^3 lattice MISE would refine to (resolution0 = 32, two upsampling steps -> 128):
        p = box_size * (index / resolution - 0.5), box_size = 1 + padding.  Returns a float32 cuda tensor
        [R+1, R+1, R+1] indexed [x, y, z] like MISE.to_dense(); every value is an evaluated one (MISE fills the
        cells it skips by propagation)."""
        capi.require_gpu()
        n = resolution + 1
        cc = torch.as_tensor(c).detach().float().reshape(1, -1).cuda().contiguous()
        ax = (1.0 + padding) * (torch.arange(n, dtype=torch.float32, device="cuda") / resolution - 0.5)
        out = torch.empty(n * n * n, dtype=torch.float32, device="cuda")
        ws = self._prepare_eval(cc, min(points_batch_size, n * n * n))
        for lo in range(0, n * n * n, points_batch_size):
            idx = torch.arange(lo, min(lo + points_batch_size, n * n * n), device="cuda")
            x = torch.stack([ax[idx // (n * n)], ax[(idx // n) % n], ax[idx % n]], dim=1).contiguous().view(1, -1, 3)
            out[lo:lo + x.shape[1]] = self._logits_prepared(x, ws)
        return out.view(n, n, n)


class ONetRestorer:
    def __init__(self, decoder, threshold=0.2, lr=1e-3):
        self.decoder, self.threshold, self.lr = decoder, float(threshold), float(lr)
        self.last_stats = None

    def optimize_points(self, opt_points, z, c, rep_weight=1., iterations=1000, printing=False, B_ref=None, normalize=True,
                        return_tensor=False):
        capi.require_gpu()
        x = opt_points.detach().float().cuda().contiguous().clone()
        cc = c.detach().float().cuda().contiguous()
        B, K, _ = x.shape
        L = capi.lib()
        P = capi.default_params(n_steps=iterations + 1, B_ref=int(B if B_ref is None else B_ref), rep_weight=float(rep_weight),
                                lr=self.lr, occ_target=self.threshold, want_stats=int(bool(printing)),
                                normalize_out=int(bool(normalize)))
        ws = self.decoder.ws.get(B, K, x.device)
        stats = torch.zeros((iterations // 100 + 1, 4), dtype=torch.float64, device=x.device) if printing else None
        capi.check(L.ifd_onet_opt(capi.ptr(self.decoder.blob), capi.ptr(cc), capi.ptr(x), None, None, B, K, ctypes.byref(P),
                                  capi.ptr(stats), capi.ptr(ws), ws.numel(), capi.stream()), "ifd_onet_opt")
        if printing:
            self.last_stats = stats.cpu().numpy()
            for j, (loss, occ, rep, sg) in enumerate(self.last_stats):
                print('iter {}, loss {:.4f}'.format(j * 100, loss))
                print('occ loss: {:.4f}, rep loss: {:.4f}\nocc value mean: {:.4f}'.format(occ, rep, sg))
        return x if return_tensor else x.cpu().numpy()
