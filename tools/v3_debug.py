import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import capi, convonet
from tests.gpu_util import dev, run_opt
z = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "convonet.npz")))
sd = {k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("sd/")}
dec = convonet.ConvONetDecoder(sd)
planes = convonet.planes_to_channels_last({k: torch.from_numpy(z["planes_nchw"][i]).cuda() for i, k in enumerate(("xz", "xy", "yz"))})
L = capi.lib()
def grad(xyz, kernel):
    x = dev(xyz); B, K, _ = x.shape
    ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(B, K), dtype=torch.uint8, device="cuda")
    g = torch.empty_like(x)
    capi.check(L.ifd_convonet_decode_bce_grad(capi.ptr(planes), capi.ptr(dec.blob), capi.ptr(x), B, K, 64, 32, 32, 5, 0.1, 0.2, B, kernel, capi.ptr(g), capi.ptr(ws), ws.numel(), capi.stream()))
    torch.cuda.synchronize(); return g.cpu().numpy()
for n in (1, 2, 3, 5):
    a, _ = run_opt(dec, planes, z["p0"], n, decode_kernel=3)
    b, _ = run_opt(dec, planes, z["p0"], n, decode_kernel=2)
    d = np.abs(a - b)
    print("steps", n, "v3 vs v2: max %.3g, count>1e-6: %d" % (d.max(), (d > 1e-6).sum()), "where:", np.argwhere(d > 1e-6)[:6].tolist())
p1, _ = run_opt(dec, planes, z["p0"], 1, decode_kernel=2)
g2, g3 = grad(p1, 2), grad(p1, 3)
d = np.abs(g2 - g3) / np.abs(g2).max()
print("gradient at step-1 positions, v3 vs v2: max rel %.3g, count>1e-5: %d" % (d.max(), (d > 1e-5).sum()), np.argwhere(d > 1e-5)[:8].tolist())
outs = [grad(p1, 3) for _ in range(5)]
print("v3 run-to-run bitwise equal:", all(np.array_equal(outs[0], o) for o in outs[1:]))
g0a, g0b = grad(z["p0"], 2), grad(z["p0"], 3)
d0 = np.abs(g0a - g0b) / np.abs(g0a).max()
print("gradient at p0, v3 vs v2: max rel %.3g" % d0.max())
bad = np.argwhere(d > 1e-5)
for (b_, i_, c_) in bad[:5]:
    print("  pt", b_, i_, "xyz", p1[b_, i_], "g2", g2[b_, i_], "g3", g3[b_, i_])
