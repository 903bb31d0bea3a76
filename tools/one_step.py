"""Run the restoration hot path a few times on one GPU and exit -- the short command ncu wraps
(profiles/README.md).  Not a benchmark: numbers printed under a profiler are never bench values."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, convonet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--decode-kernel", type=int, default=0)
ap.add_argument("--tail-ctas", type=int, default=0, help="ifd_test_hook(5, n): 1 = one-CTA tail, 2 = cluster pair, 0 = by context")
a = ap.parse_args()
case = synth.make_case(a.B, K=1024, seed=0, device="cuda")
dec = convonet.ConvONetDecoder(case.sd)
pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
L = capi.lib()
L.ifd_test_hook(5, a.tail_ctas)
P = capi.default_params(n_steps=a.iters + 1, B_ref=a.B, decode_kernel=a.decode_kernel)
ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(a.B, 1024), dtype=torch.uint8, device="cuda")
C, H, nb = dec.dims
for _ in range(a.steps):
    x = case.p0.cuda().clone()
    capi.check(L.ifd_convonet_opt(capi.ptr(pl), capi.ptr(dec.blob), capi.ptr(x), None, None, a.B, 1024, 64, C, H, nb,
                                  ctypes.byref(P), None, capi.ptr(ws), ws.numel(), capi.stream()))
torch.cuda.synchronize()
print("done", float(x.abs().max()))
