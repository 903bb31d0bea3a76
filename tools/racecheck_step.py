"""One ConvONet-Opt Adam step at config-2 size (64 x 1024) and one ONet-Mesh extraction, for compute-sanitizer:
   compute-sanitizer --tool racecheck python tools/racecheck_step.py [--mesh | --onet]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, convonet, synth  # noqa: E402
from tests.gpu_util import run_opt                # noqa: E402

if "--onet" in sys.argv:
    from ifdefense_b200 import onet as onet_mod
    case = synth.make_onet_case(4, K=1024, seed=0)
    rest = onet_mod.ONetRestorer(onet_mod.ONetDecoder(case.sd), threshold=0.2, lr=1e-3)
    out = rest.optimize_points(case.p0.cuda(), None, case.c.cuda(), rep_weight=500., iterations=1, normalize=False)
    print("onet ok", float(np.abs(out).max()))
elif "--mesh" in sys.argv:
    from ifdefense_b200 import mesh
    g = np.linspace(-0.55, 0.55, 33)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    occ = (0.42 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)) * 15
    v, f = mesh.extract_mesh(occ, 0.2, 0.1)
    pts = mesh.sample_surface_device(v, f, 256, rng=np.random.default_rng(0))
    torch.cuda.synchronize()
    print("mesh ok", tuple(v.shape), tuple(f.shape), tuple(pts.shape))
else:
    B = int(os.environ.get("RC_B", "64"))
    case = synth.make_case(B, K=1024, seed=0)
    dec = convonet.ConvONetDecoder(case.sd)
    pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
    capi.lib().ifd_test_hook(5, int(os.environ.get("RC_TAIL_CTAS", "0")))      # 1: the one-CTA form of cloud_step, 2: the cluster pair
    x, _ = run_opt(dec, pl, case.p0, 2)
    print("step ok", float(np.abs(x).max()))
