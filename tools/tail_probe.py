"""Experiment: cost of the cluster barrier in front of cloud_step_kernel's first remote store (ifd_test_hook(4, mode)):
0 none, 1 arrive.release / wait.acquire, 2 arrive.relaxed / wait.  python tools/tail_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, convonet, synth  # noqa: E402
from tools.sort_probe import run  # noqa: E402

B = 64
case = synth.make_case(B, K=1024, seed=0, device="cuda")
dec = convonet.ConvONetDecoder(case.sd)
pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
L = capi.lib()
outs = {}
for mode in (0, 1, 2, 0, 2):
    L.ifd_test_hook(4, mode)
    d, t, x = run(L, dec, pl, case.p0, B)
    outs[mode] = x
    print("bar_mode %d: decode %.1f us, tail %.1f us per Adam step" % (mode, d, t))
L.ifd_test_hook(4, 2)
print("same bits across modes:", all(np.array_equal(outs[0], outs[m]) for m in outs))
