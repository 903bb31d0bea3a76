#!/usr/bin/env python
"""bench.py -- restored clouds/sec of the ConvONet-Opt restoration hot path (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch: optimize_points on B=64 clouds x 1024 points,
201 Adam steps (--iterations=200), synthetic ModelNet40-shaped clouds, random-init weights with the
reference's parameter names (no checkpoint or data ships with the reference).  The K timed steps go through
ifd_convonet_opt_batches, which runs the loops of two consecutive batches side by side on two streams (one
launch of the loop fills 128 of the 148 SMs at B=64).  Under torchrun every rank restores its own batches (weak
scaling; clouds are independent) and the restored clouds of every step are all-gathered over NCCL -- the only
collective on the path.

Printed JSON (rank 0, one line): value = whole-job clouds/s with inputs resident in HBM (CUDA events, max
over ranks); e2e = the same through the host-buffer C-ABI call (ifd_convonet_opt_host_batches: H2D + layout
conversion + loop + D2H of every batch inside the timed region, on every rank); roofline for the dominant kernel; cpu_baseline = the oracle port timed on this box's host
cores on a bounded sample.  --impl reference times only that CPU port (the reference's op sequence on stock
PyTorch CPU; /root/reference itself does not exist on the GPU box).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, K, ITERS = 64, 1024, 200                    # configs[1]: batch=64 x 1024 pts, 200 iters (201 Adam steps)
WORKLOAD = "ConvONet-Opt batch=64x1024 pts, 200 iters (201 Adam steps), 3 planes 64^2 x 32 ch"
ALG_BYTES_PER_PT_STEP = 1560                   # 3 planes x 4 texels x 32 ch x 4 B + xyz r/w
ALG_FLOP_PER_PT_STEP = 61952                   # decoder fwd + dgrad


def ncu_traffic(kernel="convonet_decode_v4_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full summary
    (profiles/, cold-cache replay), or None."""
    try:
        txt = open(os.path.join(ROOT, "profiles", "r01_final_ncu_full_summary.txt")).read()
        sec = txt.split("==== " + kernel, 1)[1].split("====", 1)[0]
        unit = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = sec.split(key, 1)[1].split()[:2]
            tot += float(v) * unit[u]
        return tot
    except Exception:
        return None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_time(case, iterations, threads):
    """Seconds for `iterations`+1 Adam steps of the oracle port (reference op sequence, stock PyTorch CPU)."""
    import torch
    from oracle import torch_port as tp
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    tp.optimize_points(lambda p: tp.convonet_decode(case.sd, p, case.c), case.p0, rep_weight=500., iterations=iterations)
    return time.perf_counter() - t0


_JSON_OUT = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner to fd 1
    under torchrun).  Everything that goes to fd 1 from here on lands on stderr; emit() writes to the real stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(line + "\n")
    out.flush()


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric (oracle port), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from ifdefense_b200 import synth
    cores = os.cpu_count() or 1
    case = synth.make_case(B, K=K, seed=0)
    t_probe = cpu_port_time(case, 1, cores) / 2.0                      # seconds per Adam step (also warms up)
    budget = 150.0
    n_it = int(max(2, min(ITERS + 1, budget / max(t_probe * (args.steps + args.warmup), 1e-9))))
    for _ in range(args.warmup):
        cpu_port_time(case, n_it - 1, cores)
    ts = [cpu_port_time(case, n_it - 1, cores) for _ in range(args.steps)]
    per_step = float(np.sum(ts)) / (args.steps * n_it)                  # seconds per Adam step on B clouds
    full = per_step * (ITERS + 1)
    value = B / full
    sample = "B=%d x %d pts, %d of 201 Adam steps per bench step, scaled to 201" % (B, K, n_it)
    emit(json.dumps({
        "impl": "reference", "metric": "restored clouds/sec (N=1024, 200 iters)", "value": value, "unit": "clouds/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU, torch %s, %d threads" % (torch.__version__, cores)},
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-kernel", type=int, default=0, help="ifd_opt_params.decode_kernel (0 = production default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ifdefense_b200 import capi, convonet, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    capi.require_gpu()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)

    # ---- workload: NB distinct batches per rank so every step starts with planes that are not in L2
    NB = 3                                         # 3 x 100.7 MB of planes = 302 MB > 126 MB L2
    L = capi.lib()
    sd = None
    batches = []
    for j in range(NB):
        case = synth.make_case(B, K=K, seed=100 * rank + j, device="cuda", sd=sd)
        sd = case.sd
        planes_cl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
        batches.append((case, planes_cl, case.p0.cuda()))
    dec = convonet.ConvONetDecoder(sd, padding=0.1)
    C, H, nb = dec.dims
    R = batches[0][1].shape[2]
    P = capi.default_params(n_steps=ITERS + 1, B_ref=B, decode_kernel=args.decode_kernel)
    ws_bytes = L.ifd_convonet_opt_workspace_bytes(B, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    x = torch.empty((B, K, 3), dtype=torch.float32, device="cuda")
    gathered = [torch.empty_like(x) for _ in range(world)] if world > 1 else None
    stream = torch.cuda.current_stream()

    def step(j, gather=True):
        _, planes_cl, p0 = batches[j % NB]
        x.copy_(p0)
        capi.check(L.ifd_convonet_opt(capi.ptr(planes_cl), capi.ptr(dec.blob), capi.ptr(x), None, None, B, K, R, C, H, nb,
                                      ctypes.byref(P), None, capi.ptr(ws), ws_bytes, stream.cuda_stream), "ifd_convonet_opt")
        if world > 1 and gather:              # the one collective of the path: restored clouds of all ranks
            dist.all_gather(gathered, x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the timed path: K steps (batches) through ifd_convonet_opt_batches -- the loops of two consecutive batches run
    #      side by side on two streams (one launch fills 128 of the 148 SMs); every step still restores its own batch
    n_max = max(args.steps, W)
    xs = [torch.empty_like(x) for _ in range(n_max)]
    lanes = int(os.environ.get("IFD_LANES", "2"))          # experiment knob: loops side by side (default 2)
    if lanes != 2:
        L.ifd_test_hook(2, lanes)
    ws2_bytes = max(lanes, 2) * ((ws_bytes + 255) // 256 * 256)
    ws2 = torch.empty(ws2_bytes, dtype=torch.uint8, device="cuda")

    def run_steps(j0, n):
        for j in range(n):
            xs[j].copy_(batches[(j0 + j) % NB][2])
        pp = (ctypes.c_void_p * n)(*[batches[(j0 + j) % NB][1].data_ptr() for j in range(n)])
        xp = (ctypes.c_void_p * n)(*[xs[j].data_ptr() for j in range(n)])
        capi.check(L.ifd_convonet_opt_batches(n, pp, capi.ptr(dec.blob), xp, B, K, R, C, H, nb, ctypes.byref(P), capi.ptr(ws2),
                                              ws2_bytes, stream.cuda_stream), "ifd_convonet_opt_batches")
        if world > 1:                          # the one collective of the path: restored clouds of all ranks, per step
            for j in range(n):
                dist.all_gather(gathered, xs[j])

    run_steps(0, W)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    L.ifd_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(W, args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(L.ifd_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    roof = None
    e2e = None
    cpu = None
    # ---- e2e on EVERY rank: host buffers through the reference-facing host call (pinned inputs; H2D, layout conversion,
    #      loop and D2H of every batch inside the timed region), whole-job clouds/s over the slowest rank
    rest = convonet.Restorer(dec, threshold=0.2, lr=1e-3, decode_kernel=args.decode_kernel)
    host = []
    for case, _, _ in batches:
        pl = torch.stack([case.c[k] for k in ("xz", "xy", "yz")]).contiguous().pin_memory()
        host.append((pl.numpy(), case.p0.clone().pin_memory().numpy()))
    n_e2e = max(4, min(args.steps, 12))
    seq = [host[j % NB] for j in range(n_e2e)]
    out_pinned = [torch.empty((B, K, 3), dtype=torch.float32).pin_memory() for _ in range(n_e2e)]     # result buffers (host)
    out_np = [t.numpy() for t in out_pinned]
    rest.optimize_points_host_many([h[1] for h in seq[:3]], [h[0] for h in seq[:3]], rep_weight=500., iterations=ITERS, B_ref=B,
                                   out=out_np[:3])
    barrier()
    t0 = time.perf_counter()
    outs = rest.optimize_points_host_many([h[1] for h in seq], [h[0] for h in seq], rep_weight=500., iterations=ITERS, B_ref=B,
                                          out=out_np)
    if world > 1:                                  # the final gather of the restored clouds of the last batch
        dist.all_gather(gathered, torch.from_numpy(outs[-1]).cuda())
        torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    if rank == 0:
        t1 = time.perf_counter()                   # the same batches one blocking call at a time (no overlap), for reference
        for h in seq[:4]:
            rest.optimize_points_host(h[1], h[0], rep_weight=500., iterations=ITERS, B_ref=B)
        t_serial = (time.perf_counter() - t1) / 4
        h2d = int(host[0][0].nbytes + host[0][1].nbytes)
        e2e = {"value": world * B * n_e2e / t_e2e, "unit": "clouds/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(outs[-1].nbytes), "steps": n_e2e, "ms_per_step": t_e2e / n_e2e * 1e3,
               "ms_per_step_unpipelined_rank0": t_serial * 1e3,
               "api": "Restorer.optimize_points_host_many -> ifd_convonet_opt_host_batches on every rank: pinned host buffers in "
                      "and out, H2D of batch j+1 and D2H of batch j-1 overlap the loop of batch j; max over ranks"}

    # ---- rank 0: per-kernel device time for the roofline (separate pass; events bracket every launch), CPU baseline
    if rank == 0:
        L.ifd_profile_enable(1)
        n_prof = max(2, min(args.steps, 5))
        for j in range(n_prof):
            step(W + j, gather=False)          # rank 0 alone runs this pass: no collective in it
        kms = (ctypes.c_double * 4)()
        kn = (ctypes.c_longlong * 4)()
        L.ifd_profile_read(kms, kn)
        L.ifd_profile_enable(0)
        hbm_peak, peak_src = peaks()
        dec_ms = kms[0] / max(kn[0], 1)
        alg_bytes = ALG_BYTES_PER_PT_STEP * B * K                     # per decode launch (one Adam step)
        achieved = alg_bytes / (dec_ms * 1e-3) / 1e9
        total_k = sum(kms)
        roof = {"bound": "hbm", "kernel": "convonet_decode_v4_kernel (plane gather + tcgen05 ResNet-MLP fwd/dgrad)", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                "traffic_source": "profiles/r01_final_ncu_full_summary.txt (ncu --set full, bytes per launch, cold-cache replay)",
                "ms_per_launch": dec_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "fp32_tflops_achieved": ALG_FLOP_PER_PT_STEP * B * K / (dec_ms * 1e-3) / 1e12,
                "kernel_time_share": {"decode": kms[0] / total_k, "knn_repulsion_adam (cloud_step)": kms[1] / total_k,
                                      "adam (separate, legacy tail only)": kms[2] / total_k},
                "note": "planes (100 MB) are L2-resident after the first Adam step and the points of a cloud share texels, so DRAM "
                        "traffic is far below the algorithmic gather bytes; the kernel is bound by L2->SM gather latency (texels "
                        "fetched for fwd and bwd) and the tcgen05/TMEM round trip of the 30-layer chain (profiles/), not by "
                        "HBM: fp32_tflops_achieved counts the decoder's algorithmic FLOPs"}

        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            case0 = batches[0][0]
            cpu_port_time(case0, 0, cores)                              # warm-up (1 Adam step)
            n_it = 24
            t = cpu_port_time(case0, n_it - 1, cores)
            if t < 5.0:                                                 # aim for >= ~10 s of CPU work
                n_it = int(min(ITERS + 1, n_it * 10.0 / max(t, 1e-3)))
                t = cpu_port_time(case0, n_it - 1, cores)
            cpu = {"value": B / (t / n_it * (ITERS + 1)), "unit": "clouds/s", "cores": cores, "kind": "port",
                   "sample": "oracle/torch_port.py (reference op sequence, torch CPU), B=%d x %d pts, %d of 201 Adam steps in %.1f s, "
                             "scaled to 201" % (B, K, n_it, t)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(json.dumps({
            "metric": "restored clouds/sec (N=1024, 200 iters)", "value": value, "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clouds_per_gpu_per_step": B, "B_ref": B,
                       "l2": "inputs rotate over %d distinct batches per rank (%.0f MB of planes) > 126 MB L2" % (NB, NB * 100.7),
                       "concurrency": "loops of two consecutive steps run side by side on two streams (ifd_convonet_opt_batches)",
                       "parallelism": "clouds sharded over %d GPU(s), all_gather of restored clouds per step" % world},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }))


if __name__ == "__main__":
    _protect_stdout()
    main()
