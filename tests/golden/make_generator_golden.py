"""Generates tests/golden/generator.npz by running the REFERENCE's own Generator3D.generate_from_latent
(ONet/im2mesh/onet/generation.py:88-186, imported unmodified on the CPU) on the synthetic ONet: MISE is the reference's
mise.pyx cythonized from a scratch copy, marching cubes is the reference's marchingcubes.cpp compiled by oracle/Makefile
(both injected where the reference imports its compiled extensions), trimesh.Trimesh is a record of (vertices, faces).
Stored: latent code, the bias shift that puts a surface into the box, every (points, values) pair the generator evaluated
per MISE round, the dense value grid, vertices and faces.     python tests/golden/make_generator_golden.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from ifdefense_b200 import models, synth  # noqa: E402
from oracle import mcubes_ref, ref_import  # noqa: E402
from make_mise_golden import build_reference_mise  # noqa: E402


def main():
    torch.set_num_threads(1)
    ns = ref_import.load("ONet")
    _, model = ref_import.build_model(ns)
    sd = models.synthetic_state_dict("onet", 0)
    case = synth.make_onet_case(1, K=64, seed=2)
    c = case.c[:1]
    z = torch.empty(1, 0)
    model.load_state_dict(sd)
    # move the logit(0.2) level set into the box: ~30 % of a coarse lattice occupied
    ax = 1.1 * (torch.arange(17, dtype=torch.float32) / 16 - 0.5)
    pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).view(1, -1, 3)
    with torch.no_grad():
        g = model.decode(pts, z, c).logits.flatten()
    thr = float(np.log(0.2) - np.log(0.8))
    shift = np.float32(thr - float(torch.quantile(g, 0.7)))
    sd["decoder.fc_out.bias"] = sd["decoder.fc_out.bias"] + float(shift)
    model.load_state_dict(sd)

    sys.path.insert(0, ns.root)
    try:
        gen_mod = importlib.import_module("im2mesh.onet.generation")
    finally:
        sys.path.remove(ns.root)
    gen_mod.MISE = build_reference_mise()
    gen_mod.libmcubes = types.SimpleNamespace(marching_cubes=mcubes_ref.marching_cubes)
    gen_mod.trimesh = types.SimpleNamespace(
        Trimesh=lambda vertices, faces, vertex_normals=None, process=False: types.SimpleNamespace(vertices=vertices, faces=faces))
    G = gen_mod.Generator3D(model, device=torch.device("cpu"), threshold=0.2, resolution0=8, upsampling_steps=2)
    log = []
    orig = G.eval_points

    def logged(p, z_, c_=None, **kw):
        v = orig(p, z_, c_, **kw)
        log.append((p.clone(), v.clone()))
        return v
    G.eval_points = logged
    grids = []
    orig_extract = G.extract_mesh

    def extract(occ_hat, z_, c_=None, stats_dict=dict()):
        grids.append(np.array(occ_hat))
        return orig_extract(occ_hat, z_, c_, stats_dict=stats_dict)
    G.extract_mesh = extract
    mesh = G.generate_from_latent(z, c)
    out = {"c": c.numpy(), "bias_shift": shift, "value_grid": grids[0], "verts": np.asarray(mesh.vertices),
           "faces": np.asarray(mesh.faces).astype(np.int32), "rounds": np.array([len(p) for p, _ in log], dtype=np.int64),
           "points_f": np.concatenate([p.numpy() for p, _ in log]).astype(np.float32),
           "values": np.concatenate([v.numpy() for _, v in log]).astype(np.float32)}
    assert np.array_equal(out["value_grid"].astype(np.float32).astype(np.float64), out["value_grid"])
    out["value_grid"] = out["value_grid"].astype(np.float32)
    for k, v in out.items():
        print(k, getattr(v, "shape", v), getattr(v, "dtype", ""))
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)


if __name__ == "__main__":
    main()
