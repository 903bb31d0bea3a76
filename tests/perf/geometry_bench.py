"""Measurement + parity script for the classifier-side geometry kernels (BASELINE.json configs[4]: DGCNN k=20 kNN and
PointNet++ FPS / ball query at N = 1024 / 2048 / 4096, B = 32).  Test infrastructure: it times the product kernels
(CUDA events) next to the reference's own op sequence on stock PyTorch on the same GPU (full [B,N,N] distance matrix +
topk / sort, Python loop over npoint for FPS; cuBLAS rounds the K=3 products differently from the CPU's FMA chain, so
a few indices may differ from the CPU-pinned kernels -- the fraction is reported), and checks that the indices are identical.
Prints one JSON line per case.  Run on a GPU box:  python tests/perf/geometry_bench.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ifdefense_b200 import capi, defense  # noqa: E402


# The reference's op sequences on the tensor's own device (oracle/torch_port.py holds the same on the CPU):
def ref_dgcnn_knn(x, k):                       # baselines/model/dgcnn.py:7-13
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]


def ref_fps(xyz, npoint, start):               # baselines/model/pointnet2.py:53-74 (start index given instead of drawn)
    B, N, _ = xyz.shape
    dev = xyz.device
    cent = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    distance = torch.ones(B, N, device=dev) * 1e10
    far = start.to(dev).long()
    bi = torch.arange(B, dtype=torch.long, device=dev)
    for i in range(npoint):
        cent[:, i] = far
        c = xyz[bi, far, :].view(B, 1, 3)
        d = torch.sum((xyz - c) ** 2, -1)
        m = d < distance
        distance[m] = d[m]
        far = torch.max(distance, -1)[1]
    return cent


def ref_ball(radius, nsample, xyz, new_xyz):   # baselines/model/pointnet2.py:9-30,77-98
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    gi = torch.arange(N, dtype=torch.long, device=xyz.device).view(1, 1, N).repeat([B, S, 1])
    sq = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    sq += torch.sum(new_xyz ** 2, -1).view(B, S, 1)
    sq += torch.sum(xyz ** 2, -1).view(B, 1, N)
    gi[sq > radius ** 2] = N
    gi = gi.sort(dim=-1)[0][:, :, :nsample]
    first = gi[:, :, 0].view(B, S, 1).repeat([1, 1, nsample])
    mask = gi == N
    gi[mask] = first[mask]
    return gi


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    capi.require_gpu()
    B = 32
    rng = np.random.default_rng(0)
    for N in (1024, 2048, 4096):
        xyz = torch.from_numpy(rng.uniform(-1, 1, size=(B, N, 3)).astype(np.float32)).cuda()
        for C in (3, 64):
            x = xyz if C == 3 else torch.from_numpy(rng.normal(size=(B, N, C)).astype(np.float32)).cuda()
            xt = x.transpose(2, 1).contiguous()                      # DGCNN holds [B,C,N]
            ms, idx = timed(lambda: defense.dgcnn_knn(xt, 20))
            ms_ref, ref = timed(lambda: ref_dgcnn_knn(xt, 20), reps=2)
            same = float((idx == ref).float().mean())                # torch.topk on CUDA orders ties differently; random data: none
            print(json.dumps({"op": "dgcnn_knn k=20", "B": B, "N": N, "C": C, "ms": ms, "ms_torch_ops": ms_ref,
                              "speedup": ms_ref / ms, "idx_equal_frac": same, "queries_per_s": B * N / (ms * 1e-3)}))
        start = torch.zeros(B, dtype=torch.long)
        ms, f1 = timed(lambda: defense.farthest_point_sample(xyz, 512, start))
        t0 = time.perf_counter()
        ref = ref_fps(xyz, 512, start)
        torch.cuda.synchronize()
        ms_ref = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"op": "fps 512", "B": B, "N": N, "ms": ms, "ms_torch_ops": ms_ref, "speedup": ms_ref / ms,
                          "idx_equal_frac": float((f1 == ref).float().mean())}))
        new_xyz = defense.index_points(xyz, f1)
        ms, g = timed(lambda: defense.query_ball_point(0.2, 32, xyz, new_xyz))
        ms_ref, ref = timed(lambda: ref_ball(0.2, 32, xyz, new_xyz), reps=2)
        print(json.dumps({"op": "ball_query r=0.2 nsample=32", "B": B, "N": N, "S": 512, "ms": ms, "ms_torch_ops": ms_ref,
                          "speedup": ms_ref / ms, "idx_equal_frac": float((g == ref).float().mean())}))


if __name__ == "__main__":
    main()
