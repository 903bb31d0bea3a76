"""Round-2 starting point: builds tools/experiments/unet_conv3x3.cu.txt into a scratch library on the GPU box and compares the
tcgen05 implicit-GEMM convolution with F.conv2d (fp32, TF32 off) at the U-Net's shapes, then times both.
    python tools/experiments/test_unet_conv3x3.py            (needs a B200; nothing here is part of the product)"""
import ctypes
import os
import subprocess
import sys
import tempfile

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build():
    out = os.path.join(tempfile.mkdtemp(prefix="exp_conv_"), "libexp_conv.so")
    # common.cuh's helpers (fail, launch counter, set_max_dyn_smem) live in the product library: link against it
    subprocess.check_call(["nvcc", "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-shared",
                           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "if-defense_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "experiments", "unet_conv3x3.cu.txt"), "-o", out,
                           "-L", os.path.join(ROOT, "if-defense_b200"), "-lifd_b200",
                           "-Xlinker", "-rpath," + os.path.join(ROOT, "if-defense_b200")])
    return ctypes.CDLL(out)


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    L = build()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    # (B, H, C0, C1, Cout): the down path, the bottleneck and an up block with its skip connection
    for B, H, C0, C1, Cout in [(2, 64, 32, 0, 32), (2, 32, 32, 0, 64), (2, 16, 64, 0, 128), (2, 8, 128, 0, 256), (2, 16, 128, 128, 128),
                               (64, 64, 32, 0, 32)]:
        Cin = C0 + C1
        x0 = torch.randn(B, H, H, C0, device="cuda", generator=g)
        x1 = torch.randn(B, H, H, C1, device="cuda", generator=g) if C1 else None
        w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
        b = torch.randn(Cout, device="cuda", generator=g) * 0.1
        img = torch.empty(9 * (Cin // 32) * 2 * Cout * 32, device="cuda")
        out = torch.empty(B, H, H, Cout, device="cuda")
        assert L.exp_conv3x3_pack(vp(w.data_ptr()), Cout, Cin, vp(img.data_ptr()), vp(st)) == 0
        rc = L.exp_conv3x3(vp(x0.data_ptr()), vp(x1.data_ptr()) if C1 else None, vp(img.data_ptr()), vp(b.data_ptr()), vp(out.data_ptr()),
                           B, H, H, C0, C1, Cout, 1, vp(st))
        torch.cuda.synchronize()
        xin = torch.cat([x0, x1], dim=3) if C1 else x0
        want = F.relu(F.conv2d(xin.permute(0, 3, 1, 2).contiguous(), w, b, padding=1)).permute(0, 2, 3, 1)
        err = (out - want).abs().max().item()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        xn = xin.permute(0, 3, 1, 2).contiguous()
        ev[0].record()
        for _ in range(10):
            L.exp_conv3x3(vp(x0.data_ptr()), vp(x1.data_ptr()) if C1 else None, vp(img.data_ptr()), vp(b.data_ptr()), vp(out.data_ptr()),
                          B, H, H, C0, C1, Cout, 1, vp(st))
        ev[1].record()
        for _ in range(10):
            F.relu(F.conv2d(xn, w, b, padding=1))
        ev[2].record()
        torch.cuda.synchronize()
        print("B %d H %d Cin %d Cout %d: rc %d max|err| %.3g (scale %.3g)  ours %.3f ms  cuDNN fp32 %.3f ms" % (
            B, H, Cin, Cout, rc, err, want.abs().max().item(), ev[0].elapsed_time(ev[1]) / 10, ev[1].elapsed_time(ev[2]) / 10))

    # 1x1 convolution, 2x2 transposed convolution, 2x2 max pooling
    for B, H, Cin, Cout in [(2, 64, 32, 32), (2, 8, 256, 128), (2, 16, 128, 64), (2, 32, 64, 32)]:
        x = torch.randn(B, H, H, Cin, device="cuda", generator=g)
        xn = x.permute(0, 3, 1, 2).contiguous()
        w1 = torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5
        wt = torch.randn(Cin, Cout, 2, 2, device="cuda", generator=g) / Cin ** 0.5
        b = torch.randn(Cout, device="cuda", generator=g) * 0.1
        img1 = torch.empty((Cin // 32) * 2 * Cout * 32, device="cuda")
        imgt = torch.empty((Cin // 32) * 2 * 4 * Cout * 32, device="cuda")
        assert L.exp_gemm_pack(vp(w1.data_ptr()), Cout, Cin, 0, Cout, vp(img1.data_ptr()), vp(st)) == 0
        assert L.exp_gemm_pack(vp(wt.data_ptr()), 4 * Cout, Cin, 1, Cout, vp(imgt.data_ptr()), vp(st)) == 0
        o1 = torch.empty(B, H, H, Cout, device="cuda")
        ot = torch.empty(B, 2 * H, 2 * H, Cout, device="cuda")
        op = torch.empty(B, H // 2, H // 2, Cin, device="cuda")
        r1 = L.exp_conv1x1(vp(x.data_ptr()), vp(img1.data_ptr()), vp(b.data_ptr()), vp(o1.data_ptr()), B, H, H, Cin, Cout, vp(st))
        rt = L.exp_upconv2x2(vp(x.data_ptr()), vp(imgt.data_ptr()), vp(b.data_ptr()), vp(ot.data_ptr()), B, H, H, Cin, Cout, vp(st))
        rp = L.exp_maxpool2(vp(x.data_ptr()), B, H, H, Cin, vp(op.data_ptr()), vp(st))
        torch.cuda.synchronize()
        e1 = (o1 - F.conv2d(xn, w1, b).permute(0, 2, 3, 1)).abs().max().item()
        et = (ot - F.conv_transpose2d(xn, wt, b, stride=2).permute(0, 2, 3, 1)).abs().max().item()
        ep = (op - F.max_pool2d(xn, 2, 2).permute(0, 2, 3, 1)).abs().max().item()
        print("H %d Cin %d Cout %d: rc %d %d %d  1x1 err %.3g  upconv err %.3g  maxpool err %.3g" % (H, Cin, Cout, r1, rt, rp, e1, et, ep))


if __name__ == "__main__":
    sys.exit(main())
