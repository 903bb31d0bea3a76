"""Measurement script for BASELINE.json configs[3] (ONet-Mesh per cloud: 129^3 occupancy lattice + marching cubes +
1024-point surface sample, 1 GPU).  Test infrastructure: times the product path (CUDA events) stage by stage and, next to
it on the host, the reference's own marching cubes (oracle/_ref, compiled from the reference's sources) or the oracle
restatement on the same lattice, and checks that the meshes are identical.  Synthetic ONet weights and clouds (no
checkpoint ships).  Prints one JSON line.   python tests/perf/mesh_bench.py [--clouds 8]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ifdefense_b200 import capi, mesh, models, onet as onet_mod, synth  # noqa: E402
from oracle import c_oracle as co, mcubes_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=8)
    ap.add_argument("--dense", type=int, default=0, help="1: evaluate all 129^3 points instead of refining with MISE")
    a = ap.parse_args()
    capi.require_gpu()
    sd = models.synthetic_state_dict("onet", 0)
    case = synth.make_onet_case(a.clouds, K=64, seed=1)
    sd = synth.onet_with_surface(sd, case.c[:1], occupied=0.25)      # the level set passes through the box (synth.py)
    dec = onet_mod.ONetDecoder(sd)
    gen = mesh.Generator3D(dec, threshold=0.2, resolution0=32, upsampling_steps=2, padding=0.1, dense=bool(a.dense))
    rng = np.random.default_rng(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_grid = t_mc = t_samp = 0.0
    stats = []
    grids = []
    for w in range(a.clouds + 1):                      # cloud 0 twice: the first pass is the warm-up
        i = max(w - 1, 0)
        c = case.c[i:i + 1]
        u = np.concatenate([rng.random(1024)[:, None], rng.random((1024, 2))], axis=1)
        ev[0].record()
        grid = gen.value_grid(c)
        ev[1].record()
        v, f = mesh.extract_mesh(grid, 0.2, 0.1)
        ev[2].record()
        pts = mesh.sample_surface_device(v, f, 1024, uniforms=u) if f.shape[0] else None
        ev[3].record()
        torch.cuda.synchronize()
        if w == 0:
            continue
        t_grid += ev[0].elapsed_time(ev[1])
        t_mc += ev[1].elapsed_time(ev[2])
        t_samp += ev[2].elapsed_time(ev[3])
        stats.append((int(v.shape[0]), int(f.shape[0]), int(gen.points_evaluated) if not a.dense else 129 ** 3))
        grids.append((grid.cpu().numpy().astype(np.float64), v.cpu().numpy(), f.cpu().numpy()))
    # host side: the reference's marching cubes on the same lattices (np.pad + marching_cubes + transform)
    t0 = time.perf_counter()
    same = True
    for g, v, f in grids:
        if mcubes_ref.available():
            padded = np.pad(g, 1, "constant", constant_values=-1e6)
            rv, rf = mcubes_ref.marching_cubes(padded, np.log(0.2) - np.log(0.8))
            rv -= 0.5
            rv -= 1
            rv /= np.array([g.shape[0] - 1.0] * 3)
            rv = 1.1 * (rv - 0.5)
        else:
            rv, rf = co.extract_mesh(g, 0.2, 0.1)
        same = same and np.array_equal(rv, v) and np.array_equal(rf, f)
    t_ref = (time.perf_counter() - t0) * 1e3
    n = a.clouds
    per = (t_grid + t_mc + t_samp) / n
    print(json.dumps({
        "workload": "ONet-Mesh per cloud: 129^3 lattice (%s) + marching cubes + 1024 surface samples"
                    % ("dense, %d points" % 129 ** 3 if a.dense else "MISE 32 -> 128, as the reference"),
        "clouds": n, "ms_per_cloud": per, "clouds_per_s": 1e3 / per,
        "ms_lattice_eval": t_grid / n, "ms_marching_cubes": t_mc / n, "ms_surface_sample": t_samp / n,
        "verts_faces_points_evaluated": stats[:4], "mesh_equals_reference": bool(same),
        "cpu_marching_cubes_ms_per_cloud": t_ref / n,
        "cpu_kind": "reference (oracle/_ref)" if mcubes_ref.available() else "port (oracle C restatement)",
        "lattice_tflops_issued": 3 * 1312768 * float(np.mean([st[2] for st in stats])) / (t_grid / n * 1e-3) / 1e12,
        "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
