"""ctypes binding of libxvr_b200.so -- the C-ABI declared in include/xvr_b200.h.

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, an exception is
raised.  Tensors cross the boundary as raw device pointers + sizes; the CUDA stream is torch's current one.
"""

import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

from ._build import LIB

_lib = None

P = c_void_p
_SIGNATURES = {
    "xvr_abi_version": ([], c_int),
    "xvr_last_error": ([], ctypes.c_char_p),
    "xvr_launch_count": ([], c_int64),
    "xvr_volume_create": ([c_int, c_int, c_int, ctypes.POINTER(c_void_p)], c_int),
    "xvr_occupancy_create": ([c_int, c_int, c_int, ctypes.POINTER(c_void_p)], c_int),
    "xvr_volume_upload": ([P, P, P], c_int),
    "xvr_volume_destroy": ([P], c_int),
    "xvr_volume_bbox": ([P, ctypes.POINTER(c_int), P], c_int),
    "xvr_trilinear_drr_count": ([P, c_int, c_int, c_int, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_float,
                                 P, c_int, P], c_int),
    "xvr_trilinear_rays_fwd": (
        [P, P, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
         c_int, P, P, c_int, P], c_int),
    "xvr_trilinear_rays_bwd": (
        [P, P, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
         c_int, P, P, P, P, P, P, P], c_int),
    "xvr_trilinear_drr_fwd": (
        [P, P, c_int, c_int, c_int, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_float, c_int,
         c_int, P, P, c_int, P], c_int),
    "xvr_trilinear_drr_fwd_labels": (
        [P, P, c_int, c_int, c_int, P, c_int, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_float,
         c_int, c_int, P, P, c_int, P], c_int),
    "xvr_trilinear_drr_fwd_staged": (
        [P, c_int, c_int, c_int, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_float, P, P,
         P, P], c_int),
    "xvr_drr_jac_bwd": ([P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, P, P, P], c_int),
    "xvr_drr_jac_bwd_slices": ([c_int, c_int], c_int),
    "xvr_trilinear_drr_bwd_volume": (
        [P, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_float, P, c_int, c_int, c_int, P, P,
         c_int, c_int, P], c_int),
    "xvr_rays_jac_bwd": ([P, P, c_int, c_int, P, P, P, P, P], c_int),
    "xvr_ncc_fwd": ([P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, P, P, P, P, P], c_int),
    "xvr_ncc_bwd": ([P, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, P, P], c_int),
    "xvr_sobel_fwd": ([P, c_int, c_int, c_int, P, P], c_int),
    "xvr_sobel_bwd": ([P, c_int, c_int, c_int, P, P], c_int),
    "xvr_regsim_workspace_floats": ([c_int, c_int, c_int, c_int, c_int], ctypes.c_longlong),
    "xvr_regsim": ([P, P, P, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_int, c_float, c_float, c_float,
                    c_float, P, ctypes.c_longlong, P, P, P], c_int),
    "xvr_hu_stats": ([P, ctypes.c_longlong, c_float, c_float, P, P, P], c_int),
    "xvr_hu_to_density": ([P, ctypes.c_longlong, c_float, c_float, c_float, P, P, P, P], c_int),
    "xvr_reduce_rows": ([P, c_int, c_int, P, P], c_int),
    "xvr_render_epilogue": ([P, c_int, c_int, c_int, c_float, c_float, P, P, P], c_int),
    "xvr_euler_camera_fwd": ([P, P, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_float),
                              ctypes.POINTER(c_float), P, P, P], c_int),
    "xvr_euler_camera_bwd": ([P, P, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_float),
                              ctypes.POINTER(c_float), P, P, P, P], c_int),
    "xvr_pose_fwd": ([P, P, c_int, c_int, c_int, ctypes.POINTER(c_int), c_int, c_float, ctypes.POINTER(c_float),
                      ctypes.POINTER(c_float), P, P, P, P], c_int),
    "xvr_pose_bwd": ([P, P, c_int, c_int, c_int, ctypes.POINTER(c_int), c_int, c_float, ctypes.POINTER(c_float),
                      ctypes.POINTER(c_float), P, P, P, P, P], c_int),
    "xvr_reg_update": ([P, P, P, P, c_int, P, P, P, P, P, P, P, P, c_int, ctypes.POINTER(ctypes.c_double), P], c_int),
    "xvr_siddon_rays_fwd": (
        [P, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_int, P, P,
         c_int, P], c_int),
    "xvr_siddon_drr_fwd": (
        [P, P, c_int, c_int, c_int, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_float, c_float, c_int, c_int, P,
         P, c_int, P], c_int),
    "xvr_siddon_rays_bwd": (
        [P, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_int, P, P,
         P, P, P, P, c_int, P], c_int),
    "xvr_siddon_drr_bwd_volume": (
        [P, P, P, ctypes.POINTER(c_float), c_int, c_int, c_int, c_float, c_float, P, c_int, c_int, c_int, P, c_int, c_int,
         P], c_int),
    "xvr_selftest_division": ([c_int, c_int, ctypes.c_uint, P, P], c_int),
    "xvr_siddon_trace": (
        [P, P, c_int, c_int, c_int, P, P, c_int, c_int, c_float, c_float, c_int, P, P, P, c_int, P], c_int),
}

# ---- per-call kernel options (include/xvr_b200.h XVR_OPT_*).  The library itself keeps no mutable state: the
# variant travels with every call.  This host-side holder only supplies the word; tests and tuning scripts change it
# with `with options(siddon_walk=True): ...`.
_OPTION_DEFAULTS = {"ksplit": None, "siddon_walk": False, "volgrad": "brick", "siddon_tol": "production", "trim": True}
_options = dict(_OPTION_DEFAULTS)
_TOL_CODES = {"production": 0, "exact": 1, 0.5: 2, 0.25: 3, 0.125: 4}


def opts_word():
    w = 0
    if _options["ksplit"] is not None:
        if _options["ksplit"] not in (0, 1, 2, 3):
            raise ValueError("ksplit must be None (automatic) or 0..3 (log2 of the lanes per ray)")
        w |= _options["ksplit"] + 1
    if _options["siddon_walk"]:
        w |= 0x10
    if _options["volgrad"] == "gather":
        w |= 0x20
    elif _options["volgrad"] != "brick":
        raise ValueError("volgrad must be 'brick' or 'gather'")
    if not _options["trim"]:
        w |= 0x40
    w |= _TOL_CODES[_options["siddon_tol"]] << 8
    return w


class options:
    """Context manager: ``with options(siddon_walk=True, volgrad="gather"): ...``"""

    def __init__(self, **kw):
        unknown = set(kw) - set(_OPTION_DEFAULTS)
        if unknown:
            raise TypeError(f"unknown kernel option(s): {sorted(unknown)}")
        self.kw = kw

    def __enter__(self):
        self.saved = dict(_options)
        _options.update(self.kw)
        try:
            opts_word()  # validate now
        except Exception:
            self.__exit__()
            raise
        return self

    def __exit__(self, *exc):
        _options.clear()
        _options.update(self.saved)
        return False


class XvrB200Error(RuntimeError):
    pass


def lib():
    """Load the library once; fail loudly if it has not been built (python -m xvr_b200._build)."""
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise XvrB200Error(
                f"{LIB} is missing: build it with `python -m xvr_b200._build` (needs nvcc). "
                "xvr_b200 has no CPU/PyTorch fallback."
            )
        _lib = ctypes.CDLL(str(LIB))
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc, name):
    if rc != 0:
        raise XvrB200Error(f"{name} failed (code {rc}): {lib().xvr_last_error().decode()}")


_profile = None


def start_profile():
    """Record a CUDA-event pair around every C-ABI call (on torch's current stream) until stop_profile()."""
    global _profile
    _profile = {}


def stop_profile():
    """-> {entry point: [ms per call]}"""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in (prof or {}).items()}


def call(name, *args):
    if _profile is None:
        check(getattr(lib(), name)(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(getattr(lib(), name)(*args), name)
    b.record()
    _profile.setdefault(name, []).append((a, b))


def cuda_f32(t, what):
    """The kernels take contiguous fp32 CUDA tensors; anything else is an error, not a silent fallback."""
    if not t.is_cuda:
        raise XvrB200Error(f"{what} must be a CUDA tensor (xvr_b200 has no CPU path); got {t.device}")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()
