"""Helpers for the -m gpu tests: thin direct calls into the C ABI (through ifdefense_b200.capi)."""
import ctypes

import numpy as np
import torch

from ifdefense_b200 import capi, convonet


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def run_opt(dec, planes_cl, p0, n_steps, B_ref=None, normalize=0, m=None, v=None, step0=0, stats=False, **over):
    """ifd_convonet_opt on device tensors; returns (xyz, stats or None)."""
    x = dev(p0).clone()
    B, K, _ = x.shape
    C, H, nb = dec.dims
    R = planes_cl.shape[2]
    P = capi.default_params(n_steps=n_steps, B_ref=B if B_ref is None else B_ref, normalize_out=normalize, step0=step0,
                            want_stats=int(stats), **over)
    L = capi.lib()
    ws_bytes = L.ifd_convonet_opt_workspace_bytes(B, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    st = torch.zeros((max(n_steps - 1, 0) // 100 + 1, 4), dtype=torch.float64, device="cuda") if stats else None
    capi.check(L.ifd_convonet_opt(capi.ptr(planes_cl), capi.ptr(dec.blob), capi.ptr(x), capi.ptr(m), capi.ptr(v), B, K, R, C, H, nb,
                                  ctypes.byref(P), capi.ptr(st), capi.ptr(ws), ws_bytes, capi.stream()), "ifd_convonet_opt")
    torch.cuda.synchronize()
    return x.cpu().numpy(), (st.cpu().numpy() if stats else None)
