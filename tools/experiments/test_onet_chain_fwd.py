"""Round-2 starting point: builds tools/experiments/onet_chain_fwd.cu.txt on the GPU box and compares the fused 10-layer forward
chain with the product's layered path (ifd_onet_prepare + ifd_onet_decode_fwd), then times both on one MISE-sized round.
    python tools/experiments/test_onet_chain_fwd.py          (needs a B200; nothing here is part of the product)"""
import ctypes
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ifdefense_b200 import models, onet as onet_mod, synth  # noqa: E402


def build():
    out = os.path.join(tempfile.mkdtemp(prefix="exp_chain_"), "libexp_chain.so")
    subprocess.check_call(["nvcc", "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-shared",
                           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "if-defense_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "experiments", "onet_chain_fwd.cu.txt"), "-o", out,
                           "-L", os.path.join(ROOT, "if-defense_b200"), "-lifd_b200",
                           "-Xlinker", "-rpath," + os.path.join(ROOT, "if-defense_b200")])
    return ctypes.CDLL(out)


def main():
    L = build()
    vp = ctypes.c_void_p
    dec = onet_mod.ONetDecoder(models.synthetic_state_dict("onet", 0))
    case = synth.make_onet_case(1, K=64, seed=2)
    cc = case.c[:1].cuda().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    img_bytes = 2 * 10 * 8 * 2 * 256 * 32 * 4
    s_bytes = (11 * 1 * 256 * 4 + 255) // 256 * 256
    for K in (128, 1000, 100000):
        x = ((torch.rand(1, K, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(K)) - 0.5) * 1.1).contiguous()
        ws = dec._prepare_eval(cc, K)
        want = dec._logits_prepared(x, ws).clone()
        base = ws.data_ptr()
        got = torch.empty(K, dtype=torch.float32, device="cuda")
        rc = L.exp_onet_chain_fwd(vp(dec.blob.data_ptr()), vp(x.data_ptr()), 1, K, vp(base), vp(base + img_bytes),
                                  vp(base + img_bytes + s_bytes), vp(got.data_ptr()), vp(st))
        torch.cuda.synchronize()
        err = (got - want).abs().max().item()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(5):
            L.exp_onet_chain_fwd(vp(dec.blob.data_ptr()), vp(x.data_ptr()), 1, K, vp(base), vp(base + img_bytes),
                                 vp(base + img_bytes + s_bytes), vp(got.data_ptr()), vp(st))
        ev[1].record()
        for _ in range(5):
            dec._logits_prepared(x, ws)
        ev[2].record()
        torch.cuda.synchronize()
        print("K %6d: rc %d max|dlogit| %.3g (scale %.3g)  fused %.3f ms  layered %.3f ms" % (
            K, rc, err, want.abs().max().item(), ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5))


if __name__ == "__main__":
    sys.exit(main())
