"""ONet-Opt (BASELINE.json configs[0] scaled to a batch): decoder-GEMM throughput of the tcgen05 CBN-decoder chain.
Prints one JSON line: clouds/s for 201 Adam steps (extrapolated from --steps), achieved TFLOP/s of the decoder GEMMs
(algorithmic fp32 FLOPs and 3x for the 3xTF32 products actually issued) against the measured bf16 peak / 2 (the dense
TF32 rate is half the bf16 rate on this part; no TF32 GEMM peak is in MEASURED_PEAKS.json)."""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, onet, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--steps", type=int, default=21)
a = ap.parse_args()
K = 1024
case = synth.make_onet_case(a.B, K=K, seed=0, device="cuda")
dec = onet.ONetDecoder(case.sd)
L = capi.lib()
ws = dec.ws.get(a.B, K, torch.device("cuda"))
c = case.c.cuda().contiguous()
P = capi.default_params(n_steps=a.steps, B_ref=a.B)


def run():
    x = case.p0.cuda().clone()
    capi.check(L.ifd_onet_opt(capi.ptr(dec.blob), capi.ptr(c), capi.ptr(x), None, None, a.B, K, ctypes.byref(P), None,
                              capi.ptr(ws), ws.numel(), capi.stream()), "ifd_onet_opt")
    return x


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    x = run()
e1.record()
torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / reps / a.steps
flop_step = 2625536.0 * a.B * K                      # fwd + dgrad per point per step
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
tfl = flop_step / (ms_step * 1e-3) / 1e12
print(json.dumps({"workload": "ONet-Opt B=%d x %d pts, DecoderCBatchNorm hidden 256" % (a.B, K), "ms_per_adam_step": ms_step,
                  "clouds_per_s_201_steps": a.B / (ms_step * 1e-3 * 201), "decoder_tflops_algorithmic": tfl,
                  "decoder_tflops_issued_3xtf32": 3 * tfl, "tf32_peak_estimate_tflops": tf32_peak,
                  "frac_of_tf32_peak_algorithmic": tfl / tf32_peak, "frac_of_tf32_peak_issued": 3 * tfl / tf32_peak,
                  "finite": bool(torch.isfinite(x).all())}))
