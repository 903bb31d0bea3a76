"""ctypes binding of libifd_b200.so -- the C ABI declared in include/ifd_b200.h.

This is the binding a reference maintainer would add: tensors stay torch tensors, only
`data_ptr()` and the current CUDA stream cross the boundary.  Errors become RuntimeError, the reference's
only error convention (ConvONet/defense/repulsion_loss.py:37, ConvONet/defense/SOR.py:64).

There is deliberately no fallback: if the shared library is missing or the GPU is not sm_100, every call
raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libifd_b200.so")

_c_int, _c_f, _c_d, _c_sz, _vp = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_void_p


class OptParams(ctypes.Structure):
    """struct ifd_opt_params (include/ifd_b200.h)."""
    _fields_ = [
        ("n_steps", ctypes.c_int32), ("step0", ctypes.c_int32), ("B_ref", ctypes.c_int32),
        ("knn_k", ctypes.c_int32), ("normalize_out", ctypes.c_int32), ("want_stats", ctypes.c_int32),
        ("lr", _c_d), ("beta1", _c_d), ("beta2", _c_d), ("adam_eps", _c_d), ("occ_target", _c_d),
        ("rep_weight", _c_d), ("rep_radius", _c_d), ("rep_h", _c_d), ("rep_eps", _c_d), ("padding", _c_d),
        ("decode_kernel", ctypes.c_int32), ("tail_kernel", ctypes.c_int32),
    ]


class TcLinearArgs(ctypes.Structure):
    """struct ifd_tc_linear_args (include/ifd_b200.h)."""
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("n_seg", ctypes.c_int32),
        ("a_ptr", ctypes.c_void_p * 3), ("a_ld", ctypes.c_int32 * 3), ("a_width", ctypes.c_int32 * 3), ("a_relu", ctypes.c_int32 * 3),
        ("a_blocked", ctypes.c_int32 * 3), ("out_blocked", ctypes.c_int32),
        ("a_group", ctypes.c_int32 * 3),
        ("wimg", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("resid", ctypes.c_void_p), ("ld_resid", ctypes.c_int32),
        ("relu_out", ctypes.c_int32), ("out", ctypes.c_void_p), ("ld_out", ctypes.c_int32),
        ("shuffle_cout", ctypes.c_int32), ("shuffle_H", ctypes.c_int32), ("shuffle_W", ctypes.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/ifd_b200.h declares
SIGNATURES = {
    "ifd_last_error": (ctypes.c_char_p, []),
    "ifd_abi_version": (_c_int, []),
    "ifd_device_cc": (_c_int, []),
    "ifd_knn": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "ifd_knn_repulsion_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ifd_knn_repulsion": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_d, _c_d, _c_d, _vp, _vp, _vp, _vp, _vp, _c_sz, _vp]),
    "ifd_fps": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "ifd_ball_query": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_f, _c_int, _vp, _vp]),
    "ifd_sor": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_d, _vp, _vp, _vp]),
    "ifd_convonet_decoder_nfloats": (_c_sz, [_c_int, _c_int, _c_int]),
    "ifd_planes_nchw_to_cl": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp]),
    "ifd_convonet_decode_fwd": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_convonet_decode_bwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_convonet_decode_bce_grad": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _c_d, _c_int,
                                              _c_int, _vp, _vp, _c_sz, _vp]),
    "ifd_plane_bins": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_scatter_max_gather": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "ifd_scatter_mean_cl": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "ifd_grid_bins": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_convonet_grid_decode_fwd": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_convonet_grid_decode_bwd": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _vp, _vp]),
    "ifd_convonet_grid_opt": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                       ctypes.POINTER(OptParams), _vp, _vp, _c_sz, _vp]),
    "ifd_tc_packed_floats": (_c_sz, [_c_int, ctypes.POINTER(ctypes.c_int), _c_int]),
    "ifd_tc_pack": (_c_int, [_vp, _c_int, ctypes.POINTER(ctypes.c_int), _c_int, _c_int, _vp, _vp]),
    "ifd_tc_linear": (_c_int, [ctypes.POINTER(TcLinearArgs), _vp]),
    "ifd_group_max": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp]),
    "ifd_tc_blocked_floats": (_c_sz, [ctypes.c_longlong, _c_int]),
    "ifd_tc_conv3x3": (_c_int, [_vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _c_int, _c_int, _vp, _c_int, _vp]),
    "ifd_mc_workspace_bytes": (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    "ifd_mc_count": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _c_d, _vp, _c_sz,
                              ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong), _vp]),
    "ifd_mc_emit": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_d, _c_d, _c_int, _c_d, _vp, _c_sz, _vp, _vp, _vp]),
    "ifd_sample_surface_workspace_bytes": (_c_sz, [ctypes.c_longlong]),
    "ifd_cumsum_f64": (_c_int, [_vp, ctypes.c_longlong, _vp, _vp]),
    "ifd_sample_surface": (_c_int, [_vp, ctypes.c_longlong, _vp, ctypes.c_longlong, _vp, _c_int, _vp, _vp, _vp, _c_sz, _vp]),
    "ifd_mise_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ifd_mise_init": (_c_int, [_c_int, _c_int, _vp, _c_sz, _vp]),
    "ifd_mise_query": (_c_int, [_c_int, _c_int, _vp, _c_sz, _vp, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong), _vp]),
    "ifd_mise_update": (_c_int, [_c_int, _c_int, _c_d, _vp, _c_sz, _vp, _vp, ctypes.c_longlong, _vp]),
    "ifd_mise_to_dense": (_c_int, [_c_int, _c_int, _vp, _c_sz, _vp, _vp]),
    "ifd_preprocess_pc": (_c_int, [_vp, _vp, _c_int, _c_int, ctypes.c_float, _vp, _vp, _vp]),
    "ifd_opt_params_default": (None, [ctypes.POINTER(OptParams)]),
    "ifd_convonet_opt_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ifd_convonet_opt": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  ctypes.POINTER(OptParams), _vp, _vp, _c_sz, _vp]),
    "ifd_opt_tail_step": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, ctypes.POINTER(OptParams), _c_int, _vp, _vp, _vp]),
    "ifd_convonet_opt_batches_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ifd_convonet_opt_batches": (_c_int, [_c_int, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                          ctypes.POINTER(OptParams), _vp, _c_sz, _vp]),
    "ifd_convonet_opt_host": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                       ctypes.POINTER(OptParams), _vp]),
    "ifd_convonet_opt_host_batches": (_c_int, [_c_int, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                               ctypes.POINTER(OptParams)]),
    "ifd_onet_decoder_nfloats": (_c_sz, []),
    "ifd_onet_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ifd_onet_prepare": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _c_sz, _vp]),
    "ifd_onet_decode_fwd": (_c_int, [_vp, _vp, _c_int, _c_int, _vp, _vp, _c_sz, _vp]),
    "ifd_onet_decode_bwd": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _vp, _vp, _c_sz, _vp]),
    "ifd_onet_opt": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, ctypes.POINTER(OptParams), _vp, _vp, _c_sz, _vp]),
    "ifd_release_cache": (None, []),
    "ifd_launch_count": (ctypes.c_longlong, [_c_int]),
    "ifd_selftest_umma": (_c_int, [_vp, _vp, _vp, _vp]),
    "ifd_test_hook": (None, [_c_int, _c_int]),
    "ifd_profile_enable": (None, [_c_int]),
    "ifd_profile_read": (_c_int, [_vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library with typed entry points."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libifd_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` -- there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.ifd_abi_version() != 1:
            raise RuntimeError("libifd_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ifd_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what or "ifd call", rc, (msg or b"").decode()))


def default_params(**over):
    p = OptParams()
    lib().ifd_opt_params_default(ctypes.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise RuntimeError("unknown ifd_opt_params field %r" % k)
        setattr(p, k, v)
    return p


_cc_ok = set()


def require_gpu():
    """Fail loudly unless a CUDA device of compute capability 10.x is current (checked once per device: the
    property query behind ifd_device_cc costs milliseconds)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("ifdefense_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    if dev in _cc_ok:
        return
    cc = lib().ifd_device_cc()
    if cc // 10 != 10:
        raise RuntimeError("ifdefense_b200 kernels are built for sm_100a only (device reports cc %d)" % cc)
    _cc_ok.add(dev)


def use_device(device):
    """One process per GPU (the reference's model: torch.distributed.launch --nproc_per_node): the C library launches on
    the CURRENT CUDA device and its current stream.  Objects built with an explicit `device='cuda:N'` make N current
    here, so that kernels, streams and cached buffers all live on the device that holds the tensors."""
    import torch
    device = torch.device(device)
    if device.type == "cuda" and torch.cuda.is_available():
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx != torch.cuda.current_device():
            torch.cuda.set_device(idx)
    return device


def ptr(t):
    """Device/host pointer of a contiguous tensor (or None)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("tensor passed to the C ABI must be contiguous")
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
