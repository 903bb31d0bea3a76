"""Per-kernel accuracy probe: decode+BCE gradient of kernels 1/2/3 against a float64 evaluation of the oracle."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ifdefense_b200 import capi, convonet, synth
from oracle import torch_port as tp
torch.set_num_threads(8)
case = synth.make_case(4, K=1024, seed=0, device="cuda")
dec = convonet.ConvONetDecoder(case.sd)
pl = convonet.planes_to_channels_last({k: v.cuda() for k, v in case.c.items()})
def oracle(dtype):
    sd = {k: v.to(dtype) for k, v in case.sd.items()}
    c = {k: v.to(dtype) for k, v in case.c.items()}
    p = case.p0.to(dtype).requires_grad_()
    def lin(name, x): return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])
    cc = 0
    for plane in ("xz", "xy", "yz"):
        xy = tp.normalize_coordinate(p.clone(), padding=0.1, plane=plane)
        cc = cc + F.grid_sample(c[plane], 2.0 * xy[:, :, None] - 1.0, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1)
    cc = cc.transpose(1, 2)
    net = lin("decoder.fc_p", p)
    for i in range(5):
        net = net + lin("decoder.fc_c.%d" % i, cc)
        net = net + lin("decoder.blocks.%d.fc_1" % i, F.relu(lin("decoder.blocks.%d.fc_0" % i, F.relu(net))))
    lg = lin("decoder.fc_out", F.relu(net)).squeeze(-1)
    (F.binary_cross_entropy_with_logits(lg, torch.full_like(lg, 0.2), reduction="none").mean() * p.shape[1]).backward()
    return p.grad.double().numpy()
g64, g32 = oracle(torch.float64), oracle(torch.float32)
scale = np.abs(g64).max()
print("torch fp32 vs fp64: max %.3g  median %.3g (relative to max|g|=%.3g)" % (np.abs(g32 - g64).max() / scale, np.median(np.abs(g32 - g64)) / scale, scale))
L = capi.lib()
x = case.p0.cuda().contiguous()
ws = torch.empty(L.ifd_convonet_opt_workspace_bytes(4, 1024), dtype=torch.uint8, device="cuda")
for kernel in (1, 2, 3):
    g = torch.empty_like(x)
    capi.check(L.ifd_convonet_decode_bce_grad(capi.ptr(pl), capi.ptr(dec.blob), capi.ptr(x), 4, 1024, 64, 32, 32, 5, 0.1, 0.2, 4, kernel,
                                              capi.ptr(g), capi.ptr(ws), ws.numel(), capi.stream()))
    torch.cuda.synchronize()
    d = np.abs(g.cpu().numpy().astype(np.float64) - g64) / scale
    big = (d > 1e-4).sum()
    print("kernel %d vs fp64: max %.3g  p99.9 %.3g  median %.3g  entries>1e-4: %d" % (kernel, d.max(), np.quantile(d, 0.999), np.median(d), big))
