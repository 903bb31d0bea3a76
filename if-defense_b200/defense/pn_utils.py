"""Index-producing geometry ops with the reference's names and argument meaning.

  knn_point              ConvONet/defense/pn_utils.py:64-83
  index_points           ConvONet/defense/pn_utils.py:6-23
  farthest_point_sample  ConvONet/defense/pn_utils.py:26-48, baselines/model/pointnet2.py:53-74
  fps_points             ConvONet/defense/pn_utils.py:51-61
  query_ball_point       baselines/model/pointnet2.py:77-98
  dgcnn_knn              baselines/model/dgcnn.py:7-13

Indices are returned as int64 tensors like the reference (the kernels produce int32).
"""
import numpy as np
import torch

from .. import capi


def _f32c(x):
    return x.detach().float().contiguous()


def index_points(points, idx):
    """points [B,N,C], idx [B,...] long -> [B,...,C] (plain torch gather; differentiable like the reference)."""
    B = points.shape[0]
    view = [B] + [1] * (idx.dim() - 1)
    b = torch.arange(B, dtype=torch.long, device=points.device).view(view).expand_as(idx)
    return points[b, idx, :]


def knn_point(k, points):
    """kNN idx [B,K,k] (self "assumed" in column 0 and dropped, exactly as the reference does)."""
    capi.require_gpu()
    pc = _f32c(points)
    B, K, C = pc.shape
    idx = torch.empty((B, K, k), dtype=torch.int32, device=pc.device)
    capi.check(capi.lib().ifd_knn(capi.ptr(pc), B, K, C, k, 1, capi.ptr(idx), None, capi.stream()), "ifd_knn")
    return idx.long()


def dgcnn_knn(x, k):
    """DGCNN knn: x is [B,C,N] (channels first, as the model holds it); self included; -> [B,N,k]."""
    capi.require_gpu()
    xt = _f32c(x.transpose(2, 1))
    B, N, C = xt.shape
    idx = torch.empty((B, N, k), dtype=torch.int32, device=xt.device)
    capi.check(capi.lib().ifd_knn(capi.ptr(xt), B, N, C, k, 0, capi.ptr(idx), None, capi.stream()), "ifd_knn")
    return idx.long()


def farthest_point_sample(xyz, num_point, start=None):
    """FPS centroids [B,num_point].  The reference seeds with torch.randint (pn_utils.py:39); pass `start`
    ([B] long) to make the draw explicit, otherwise it is drawn here the same way."""
    capi.require_gpu()
    pc = _f32c(xyz)
    B, N, _ = pc.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    s32 = start.to(device=pc.device, dtype=torch.int32).contiguous()
    out = torch.empty((B, num_point), dtype=torch.int32, device=pc.device)
    capi.check(capi.lib().ifd_fps(capi.ptr(pc), B, N, num_point, capi.ptr(s32), capi.ptr(out), capi.stream()), "ifd_fps")
    return out.long()


def fps_points(xyz, num_point, start=None):
    return index_points(xyz, farthest_point_sample(xyz, num_point, start))


def query_ball_point(radius, nsample, xyz, new_xyz):
    """group_idx [B,S,nsample]: first nsample ascending indices within `radius`, padded with the first."""
    capi.require_gpu()
    a, b = _f32c(xyz), _f32c(new_xyz)
    B, N, _ = a.shape
    S = b.shape[1]
    r2 = float(np.float32(radius ** 2))          # torch rounds the Python scalar to the tensor dtype
    out = torch.empty((B, S, nsample), dtype=torch.int32, device=a.device)
    capi.check(capi.lib().ifd_ball_query(capi.ptr(a), capi.ptr(b), B, N, S, r2, nsample, capi.ptr(out), capi.stream()),
               "ifd_ball_query")
    return out.long()
