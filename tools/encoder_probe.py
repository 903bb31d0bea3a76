"""Where the time of LocalPoolPointnet.forward goes at B = 192 (the reference batch): stage timers + top CUDA kernels.
python tools/encoder_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ifdefense_b200 import models, synth  # noqa: E402


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    model = models.build_convonet()
    model.load_state_dict(models.synthetic_state_dict("convonet", 0))
    enc = model.encoder.cuda().eval()
    p = (torch.rand(B, 600, 3, device="cuda") - 0.5) * 0.9
    with torch.no_grad():
        print("encoder total ms:", timed(lambda: enc(p)))
        plane_cl = torch.randn(B, 64, 64, 32, device="cuda").permute(0, 3, 1, 2)
        plane_nchw = plane_cl.contiguous()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        print("unet fp32, channels_last in, one plane ms:", timed(lambda: enc.unet(plane_cl)))
        print("unet fp32, NCHW in, one plane ms:", timed(lambda: enc.unet(plane_nchw)))
        torch.use_deterministic_algorithms(True, warn_only=True)
        print("unet fp32 deterministic, channels_last ms:", timed(lambda: enc.unet(plane_cl)))
        print("unet fp32 deterministic, NCHW ms:", timed(lambda: enc.unet(plane_nchw)))
        torch.use_deterministic_algorithms(False)
        torch.backends.cudnn.benchmark = True
        print("unet fp32 benchmark, channels_last ms:", timed(lambda: enc.unet(plane_cl)))
        print("unet fp32 benchmark, NCHW ms:", timed(lambda: enc.unet(plane_nchw)))
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.allow_tf32 = True
        print("unet TF32 (info only), channels_last ms:", timed(lambda: enc.unet(plane_cl)))
        torch.backends.cudnn.allow_tf32 = False
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            enc(p)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))


if __name__ == "__main__":
    main()
