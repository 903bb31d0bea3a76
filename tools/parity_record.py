#!/usr/bin/env python
"""Record how far the restoration loop on the B200 sits from the oracle, next to the oracle's own noise floor.

  python tools/parity_record.py [--out gpurun_out/r02_parity_record.json] [--quick]

For ConvONet-Opt at BASELINE.json configs[1] (B = 64 x 1024 points) and on the K = 256 golden fixture, and for ONet-Opt
(B = 64 at 20 steps, B = 8 at 201 steps), the |d xyz| distribution (median, p99, p99.9, max, fraction <= 1e-4 / 1e-5) after
1, 2, 10, 20, 100 and 201 Adam steps (raw coordinates, before normalize_batch_pc) for

  * oracle with all host threads  vs  oracle with half of them   -- the reference's run-to-run reproducibility,
  * oracle  vs  oracle started one ulp away (1 % of the input coordinates moved to the next float)  -- the sensitivity of
    the 201-step trajectory itself: what ANY implementation that is not bit-identical to stock PyTorch CPU must expect,
  * GPU fp32 kernels (decode_kernel = 2)  vs  oracle,
  * GPU production default (tcgen05, 3xTF32)  vs  oracle,
  * GPU default vs GPU fp32.

TEST INFRASTRUCTURE: uses oracle/ as the checker.  Needs a B200 (the GPU side goes through the C ABI).  The table in
DESIGN.md section 5 and the thresholds of tests/test_gpu_parity_full.py come from this script's output.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ifdefense_b200 import capi, convonet, onet as onet_mod, synth       # noqa: E402
from oracle import torch_port as tp                                      # noqa: E402
from tests.gpu_util import run_opt                                       # noqa: E402

STEPS = (1, 2, 10, 20, 100, 201)


def dist_row(a, b):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).ravel()
    return {"median": float(np.median(d)), "p99": float(np.quantile(d, 0.99)), "p999": float(np.quantile(d, 0.999)),
            "max": float(d.max()), "frac_le_1e-4": float((d <= 1e-4).mean()), "frac_le_1e-5": float((d <= 1e-5).mean()),
            "frac_le_1e-6": float((d <= 1e-6).mean())}


def oracle_trace(decode_fn, p0, n_steps, threads):
    torch.set_num_threads(threads)
    tr = {}
    t0 = time.perf_counter()
    tp.optimize_points(decode_fn, p0, rep_weight=500., iterations=n_steps - 1, normalize=False, trace=tr,
                       trace_steps=[s - 1 for s in STEPS if s <= n_steps])
    return {s: tr["xyz"][s - 1] for s in STEPS if s <= n_steps}, time.perf_counter() - t0


def convonet_block(B, K, seed, n_steps, cores, noise_floor=True):
    case = synth.make_case(B, K=K, seed=seed)
    dec = convonet.ConvONetDecoder(case.sd, padding=0.1)
    planes = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    fn = lambda p: tp.convonet_decode(case.sd, p, case.c)
    ref, t_ref = oracle_trace(fn, case.p0, n_steps, cores)
    out = {"B": B, "K": K, "oracle_threads": cores, "oracle_seconds": t_ref, "rows": {}}
    alt = ulp = None
    if noise_floor:
        alt, t_alt = oracle_trace(fn, case.p0, n_steps, max(1, cores // 2))
        out["oracle_alt_threads"] = max(1, cores // 2)
        out["oracle_alt_seconds"] = t_alt
        g = torch.Generator().manual_seed(11)
        sel = torch.rand(case.p0.shape, generator=g) < 0.01
        p1 = torch.where(sel, torch.nextafter(case.p0, torch.full_like(case.p0, 10.0)), case.p0)
        ulp, _ = oracle_trace(fn, p1, n_steps, cores)
        out["one_ulp_perturbed_fraction"] = float(sel.float().mean())
    for s in STEPS:
        if s > n_steps:
            continue
        g2, _ = run_opt(dec, planes, case.p0, s, decode_kernel=2)
        g0, _ = run_opt(dec, planes, case.p0, s)
        row = {"gpu_fp32_vs_oracle": dist_row(g2, ref[s]), "gpu_default_vs_oracle": dist_row(g0, ref[s]),
               "gpu_default_vs_gpu_fp32": dist_row(g0, g2)}
        if alt is not None:
            row["oracle_vs_oracle_half_threads"] = dist_row(alt[s], ref[s])
            row["oracle_vs_oracle_1ulp_start"] = dist_row(ulp[s], ref[s])
        out["rows"][str(s)] = row
    return out


def fixture_block():
    conv = dict(np.load(os.path.join(ROOT, "tests", "golden", "convonet.npz")))
    sd = {k[3:]: torch.from_numpy(v) for k, v in conv.items() if k.startswith("sd/")}
    planes_nchw = {k: torch.from_numpy(conv["planes_nchw"][i]) for i, k in enumerate(("xz", "xy", "yz"))}
    dec = convonet.ConvONetDecoder(sd, padding=0.1)
    planes = convonet.planes_to_channels_last({k: v.cuda() for k, v in planes_nchw.items()})
    out = {"B": int(conv["p0"].shape[0]), "K": int(conv["p0"].shape[1]),
           "reference": "tests/golden/convonet.npz (generated from the reference's own classes)", "rows": {}}
    for s in (1, 2, 10, 20, 201):
        ref = conv["final_201_raw"] if s == 201 else conv["trace/xyz_%d" % (s - 1)]
        g2, _ = run_opt(dec, planes, conv["p0"], s, decode_kernel=2)
        g0, _ = run_opt(dec, planes, conv["p0"], s)
        out["rows"][str(s)] = {"gpu_fp32_vs_reference": dist_row(g2, ref), "gpu_default_vs_reference": dist_row(g0, ref),
                               "gpu_default_vs_gpu_fp32": dist_row(g0, g2)}
    return out


def onet_block(B, K, seed, n_steps, cores):
    case = synth.make_onet_case(B, K=K, seed=seed)
    dec = onet_mod.ONetDecoder(case.sd)
    rest = onet_mod.ONetRestorer(dec, threshold=0.2, lr=1e-3)
    fn = lambda p: tp.onet_decode(case.sd, p, case.c)
    ref, t_ref = oracle_trace(fn, case.p0, n_steps, cores)
    out = {"B": B, "K": K, "oracle_threads": cores, "oracle_seconds": t_ref, "rows": {}}
    for s in STEPS:
        if s > n_steps:
            continue
        g = rest.optimize_points(case.p0.cuda(), None, case.c.cuda(), rep_weight=500., iterations=s - 1, normalize=False)
        out["rows"][str(s)] = {"gpu_default_vs_oracle": dist_row(g, ref[s])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity_record.json"))
    ap.add_argument("--quick", action="store_true", help="20 steps only, no noise floor (smoke run of the script)")
    args = ap.parse_args()
    capi.require_gpu()
    cores = os.cpu_count() or 1
    n = 20 if args.quick else 201
    res = {"host_cores": cores, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
           "what": "|xyz_gpu - xyz_oracle| after n Adam steps, raw coordinates (box scale ~1, point spacing ~0.03)"}
    res["convonet_fixture_2x256"] = fixture_block()
    res["convonet_config2_64x1024"] = convonet_block(64, 1024, 0, n, cores, noise_floor=not args.quick)
    res["onet_64x1024_20steps"] = onet_block(64, 1024, 0, 20 if not args.quick else 2, cores)
    if not args.quick:
        res["onet_8x1024_201steps"] = onet_block(8, 1024, 1, 201, cores)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    for blk, v in res.items():
        if isinstance(v, dict) and "rows" in v:
            print("==", blk)
            for s, row in v["rows"].items():
                for name, r in row.items():
                    print("  steps %4s  %-34s median %.2e  p99 %.2e  max %.2e  <=1e-4: %.5f" % (s, name, r["median"], r["p99"], r["max"], r["frac_le_1e-4"]))


if __name__ == "__main__":
    main()
