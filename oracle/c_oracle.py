"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/c/libifd_oracle.so (see ifd_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libifd_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def knn(x, k, drop_first=1, want_keys=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    idx = np.empty((B, N, k), dtype=np.int32)
    keys = np.empty((B, N, k), dtype=np.float32) if want_keys else None
    lib().ifdo_knn(_p(x), B, N, C, k, drop_first, _p(idx), _p(keys) if want_keys else None)
    return (idx, keys) if want_keys else idx


def fps(xyz, npoint, start):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    start = np.ascontiguousarray(start, dtype=np.int32)
    out = np.empty((B, npoint), dtype=np.int32)
    lib().ifdo_fps(_p(xyz), B, N, npoint, _p(start), _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), dtype=np.int32)
    lib().ifdo_ball_query(_p(xyz), _p(new_xyz), B, N, S, ctypes.c_float(np.float32(radius ** 2)), nsample, _p(out))
    return out


def sor(xyz, k=2, alpha=1.1):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, K, _ = xyz.shape
    keep = np.empty((B, K), dtype=np.uint8)
    val = np.empty((B, K), dtype=np.float64)
    lib().ifdo_sor(_p(xyz), B, K, k, ctypes.c_double(alpha), _p(keep), _p(val))
    return keep.astype(bool), val
