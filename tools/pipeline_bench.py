"""End-to-end defense throughput of the library driver (BASELINE.json configs[2]-like: a file of clouds in reference
batches of 192, SOR -> preprocess -> encode -> init -> 201 Adam steps -> normalise), with a per-stage breakdown, for the
device pre-processing path and the per-cloud numpy path.  Synthetic clouds and weights.  python tools/pipeline_bench.py"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import driver, models, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=384)
    ap.add_argument("--batch", type=int, default=192)
    a = ap.parse_args()
    if os.environ.get("IFD_JAC"):                 # experiment knob (see bench.py)
        from ifdefense_b200 import capi
        capi.lib().ifd_test_hook(7, int(os.environ["IFD_JAC"]))
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    base = synth.clouds(32)
    pc = np.concatenate([base] * ((a.clouds + 31) // 32))[:a.clouds]
    res = {}
    for dev_path in (True, False):
        d = driver.Defender(model, driver.Args(batch_size=a.batch, iterations=200, device_preprocess=dev_path))
        # warm-up: enough batches that every loop stream, lane graph and allocator pool of the pipeline has been used once
        d.defend_point_cloud(pc[:min(len(pc), 4 * a.batch)], rng=np.random.default_rng(0), gen=torch.Generator().manual_seed(0))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.defend_point_cloud(pc, rng=np.random.default_rng(0), gen=torch.Generator().manual_seed(0))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # stage breakdown of one batch: second of two passes (the first one after the pipelined run pays one-off allocator work on
        # the main stream -- the pipelined run allocated on its own streams)
        for _pass in range(2):
          st = {}
          x = pc[:a.batch]
          t0 = time.perf_counter()
          if dev_path:
              sel, pts = d.prepare_batch_device(x, np.random.default_rng(0), torch.Generator().manual_seed(0))
          else:
              pcs = d.sor_process(x)
              proc = [driver.preprocess_pc(p, num_points=600, padding_scale=0.9, rng=np.random.default_rng(0)) for p in pcs]
              sel = torch.from_numpy(np.stack([s for _, s in proc])).float().cuda()
              pts = driver.init_points([p for p, _ in proc], 1024, 0.01, 0.9, torch.Generator().manual_seed(0))
          torch.cuda.synchronize()
          st["prepare_ms"] = 1e3 * (time.perf_counter() - t0)
          t0 = time.perf_counter()
          with torch.no_grad():
              c = d.model.encode_inputs(sel)
          torch.cuda.synchronize()
          st["encode_ms"] = 1e3 * (time.perf_counter() - t0)
          t0 = time.perf_counter()
          d.restorer.optimize_points(pts, None, c, rep_weight=500., iterations=200)
          torch.cuda.synchronize()
          st["optimize_ms"] = 1e3 * (time.perf_counter() - t0)
        res["device_preprocess" if dev_path else "numpy_preprocess"] = {"clouds_per_s": a.clouds / dt, "seconds": dt, "one_batch": st}
    print(json.dumps({"workload": "defend_point_cloud: %d clouds x 1024 pts, batches of %d, SOR on, 201 Adam steps" % (a.clouds, a.batch),
                      "gpu": torch.cuda.get_device_name(0), **res}))


if __name__ == "__main__":
    main()
