"""ctypes binding of libpof.so (include/pof.h).  No CPU fallback, by design.

`lib()` loads the in-tree shared object (building it with nvcc first if the
sources are newer and nvcc is present) and declares every entry point's
signature.  Every compute wrapper in this package goes through `check()`, which
turns a non-zero status into `RuntimeError(pof_last_error())` — the reference's
error convention is plain Python exceptions.
"""
import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POF_LIB") or os.path.join(_PKG_DIR, "libpof.so")      # POF_LIB: a tuning variant (build.build_variant)
ABI_VERSION = 2

_lock = threading.Lock()
_lib = None

c_int = ctypes.c_int
c_size_t = ctypes.c_size_t
c_double = ctypes.c_double
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/pof.h one to one
SIGNATURES = {
    "pof_abi_version": (c_int, []),
    "pof_last_error": (ctypes.c_char_p, []),
    "pof_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "pof_cutout_ws_bytes": (c_size_t, [c_int]),
    "pof_cutout_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_double, c_double, c_double, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pof_spaam_gate_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pof_spaam_gate_bwd_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pof_spaam_gate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pof_act_fwd": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pof_conv_first_fwd": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_float,
                                   c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pof_conv_tc_fwd": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pof_conv_tc_f16_fwd": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pof_head_fwd": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_void_p]),
    "pof_patch_corr_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pof_patch_corr_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pof_cutout_original_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_double, c_double, c_double,
                                        c_int, c_int, c_void_p, c_void_p]),
    "pof_polar_grid_fwd": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_int, c_void_p, c_void_p]),
    "pof_bn_act_stats": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "pof_conv_first_wgrad": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pof_bn_act_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_float, c_float, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pof_bn_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pof_nms_ws_bytes": (c_size_t, [c_int, c_int]),
    "pof_nms_centers": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
}


def lib():
    """Return the loaded library; raise if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get("POF_LIB"):          # a tuning variant is loaded as it is
            from . import build as _build

            force = os.environ.get("POF_REBUILD") == "1"
            missing = not os.path.isfile(LIB_PATH)
            if force or missing or _build._stale():
                try:
                    _build.build(force=force)
                except RuntimeError:
                    if force or missing:               # nothing to load: fail loudly (there is no CPU fallback)
                        raise
                    import warnings                    # no nvcc on this box: say that the sources are newer than the binary

                    warnings.warn("libpof.so is older than its sources and could not be rebuilt here; loading the stale binary")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # noqa: PERF203
            raise RuntimeError(
                "libpof.so could not be loaded from %s (%s). The CUDA extension is the only "
                "implementation of this path: build it with `python -m planar_optical_flow_b200.build`." % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:
                raise RuntimeError("libpof.so is missing symbol %s (stale build?)" % name)
            fn.restype = res
            fn.argtypes = args
        got = handle.pof_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError("libpof.so ABI version %d, host code expects %d" % (got, ABI_VERSION))
        _lib = handle
    return _lib


def last_error():
    msg = lib().pof_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what):
    if status != 0:
        raise RuntimeError("%s failed (status %d): %s" % (what, status, last_error()))


def device_info():
    sm, major, minor = c_int(0), c_int(0), c_int(0)
    check(lib().pof_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "pof_device_info")
    return sm.value, major.value, minor.value


def require_cuda_tensor(t, name, dtype=None):
    """The product path only runs on device memory."""
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: planar_optical_flow_b200 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def current_stream_ptr(device):
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
