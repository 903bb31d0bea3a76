"""TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/_ref/libmcubes_ref.so -- the reference's own
`libmcubes.marching_cubes(volume, isovalue)` (ONet/im2mesh/utils/libmcubes/pywrapper.cpp:90-127 over
marchingcubes.h:23-189), compiled from the reference's sources by oracle/Makefile."""
import ctypes
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmcubes_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        _lib.refmc_run.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_double,
                                   ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
        _lib.refmc_run.restype = None
        _lib.refmc_fetch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.refmc_fetch.restype = None
    return _lib


def marching_cubes(volume, isovalue):
    """-> (verts [V,3] float64, faces [F,3] int64), exactly what `libmcubes.marching_cubes` returns (mcubes.pyx:20-25)."""
    v = np.ascontiguousarray(volume, dtype=np.float64)
    assert v.ndim == 3
    nv, ni = ctypes.c_long(), ctypes.c_long()
    lib().refmc_run(v.ctypes.data, v.shape[0], v.shape[1], v.shape[2], float(isovalue), ctypes.byref(nv), ctypes.byref(ni))
    verts = np.empty(nv.value, dtype=np.float64)
    faces = np.empty(ni.value, dtype=np.int64)
    lib().refmc_fetch(verts.ctypes.data, faces.ctypes.data)
    return verts.reshape(-1, 3), faces.reshape(-1, 3)
