"""Quick device-time numbers for the pieces around the ConvONet loop (not a benchmark line): encoders at B = 192, ONet-Opt
decoder per Adam step at B = 64.  python tools/quick_perf.py"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, models, onet as onet_mod, synth  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


L = capi.lib()
m = models.build_convonet()
m.load_state_dict(models.synthetic_state_dict("convonet", 0))
m = m.cuda().eval()
x = (torch.rand(192, 600, 3, device="cuda") - 0.5) * 0.9
planes_in = torch.randn(576, 64, 64, 32, device="cuda")
for cl in (2,):
    L.ifd_test_hook(6, cl)
    with torch.no_grad():
        print("cluster %d: ConvONet encoder, 192 clouds x 600 pts: %.2f ms" % (cl, timed(lambda: m.encode_inputs(x))))
        print("cluster %d:   of which U-Net on 576 planes: %.2f ms" % (cl, timed(lambda: m.encoder.unet.forward_cl(planes_in))))
m2 = models.build_onet()
m2.load_state_dict(models.synthetic_state_dict("onet", 0))
m2 = m2.cuda().eval()
x2 = (torch.rand(192, 300, 3, device="cuda") - 0.5) * 0.9
with torch.no_grad():
    print("ONet encoder, 192 clouds x 300 pts: %.2f ms" % timed(lambda: m2.encode_inputs(x2)))
case = synth.make_onet_case(64, K=1024, seed=0, device="cuda")
rest = onet_mod.ONetRestorer(onet_mod.ONetDecoder(case.sd))
p0, c = case.p0.cuda(), case.c.cuda()
t = timed(lambda: rest.optimize_points(p0, None, c, rep_weight=500., iterations=19, return_tensor=True), n=3)
print("ONet-Opt 64 x 1024: %.3f ms per Adam step" % (t / 20))
