"""Experiment: does a one-time spatial (Morton) ordering of each cloud's points make the per-step kernels faster?  decode
tiles become spatially compact (texel reuse in L1/L2), kNN warps walk the same grid rows.  Prints per-kernel device time
(library events) for the unsorted and sorted order of the SAME clouds.  python tools/sort_probe.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, convonet, synth  # noqa: E402


def morton_order(p, bits=5):
    lo, hi = p.min(0), p.max(0)
    q = np.clip(((p - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(len(p), dtype=np.int64)
    for b in range(bits):
        for ax in range(3):
            code |= ((q[:, ax] >> b) & 1) << (3 * b + ax)
    return np.argsort(code, kind="stable")


def run(L, dec, pl, p0, B, iters=100):
    P = capi.default_params(n_steps=iters, B_ref=B)
    ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(B, 1024), dtype=torch.uint8, device="cuda")
    C, H, nb = dec.dims
    out = {}
    for rep in range(2):
        x = p0.cuda().clone()
        L.ifd_profile_enable(1 if rep == 1 else 0)
        capi.check(L.ifd_convonet_opt(capi.ptr(pl), capi.ptr(dec.blob), capi.ptr(x), None, None, B, 1024, 64, C, H, nb,
                                      ctypes.byref(P), None, capi.ptr(ws), ws.numel(), capi.stream()))
        torch.cuda.synchronize()
    kms = (ctypes.c_double * 4)()
    kn = (ctypes.c_longlong * 4)()
    L.ifd_profile_read(kms, kn)
    L.ifd_profile_enable(0)
    return kms[0] / max(kn[0], 1) * 1e3, kms[1] / max(kn[1], 1) * 1e3, x.cpu().numpy()


def main():
    B = 64
    case = synth.make_case(B, K=1024, seed=0, device="cuda")
    dec = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    L = capi.lib()
    p0 = case.p0.numpy()
    perm = np.stack([morton_order(p0[b]) for b in range(B)])
    ps = torch.from_numpy(np.stack([p0[b][perm[b]] for b in range(B)]))
    d0, t0, x0 = run(L, dec, pl, case.p0, B)
    d1, t1, x1 = run(L, dec, pl, ps, B)
    inv = np.stack([np.argsort(perm[b]) for b in range(B)])
    x1u = np.stack([x1[b][inv[b]] for b in range(B)])
    dd = np.abs(x1u - x0)
    print("unsorted: decode %.1f us, tail %.1f us per Adam step" % (d0, t0))
    print("morton  : decode %.1f us, tail %.1f us per Adam step" % (d1, t1))
    print("100-step result, sorted vs unsorted order: median |d| %.2e, max %.2e, within 1e-5: %.5f" % (np.median(dd), dd.max(), (dd < 1e-5).mean()))


if __name__ == "__main__":
    main()
