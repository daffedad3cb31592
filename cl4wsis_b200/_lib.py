"""ctypes binding of ``libcl4wsis_b200.so`` (the C ABI declared in ``include/cl4wsis_b200.h``).

There is no CPU or PyTorch fallback: if the library is missing, or a tensor is not on a
CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CL4_LIB") or os.path.join(_HERE, "libcl4wsis_b200.so")  # CL4_LIB: A/B variant builds

CL4_OK, CL4_EINVAL, CL4_EUNSUPPORTED, CL4_ECUDA, CL4_ESCRATCH = 0, -1, -2, -3, -4
MAX_DILATIONS = 8

_vp = ctypes.c_void_p
_int = ctypes.c_int
_sz = ctypes.c_size_t
_flt = ctypes.c_float

# name -> (restype, argtypes); mirrors include/cl4wsis_b200.h one to one
SIGNATURES = {
    "cl4_abi_version": (_int, []),
    "cl4_last_error": (ctypes.c_char_p, []),
    "cl4_local_affinity": (_int, [_vp, _vp, _int, _int, _int, ctypes.POINTER(_int), _int, _int, _vp]),
    "cl4_local_stdev": (_int, [_vp, _vp, _int, _int, _int, ctypes.POINTER(_int), _int, _vp]),
    "cl4_resize_bilinear_ac": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "cl4_pamr_weights": (_int, [_vp, _vp, _int, _int, _int, _int, ctypes.POINTER(_int), _int, _vp]),
    "cl4_pamr_sweep": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, ctypes.POINTER(_int), _int, _vp]),
    "cl4_pamr_scratch_bytes": (_sz, [_int] * 7),
    "cl4_pamr_forward": (_int, [_vp, _vp, _vp, _vp, _sz, _int, _int, _int, _int, _int, ctypes.POINTER(_int), _int,
                                _int, _vp]),
    "cl4_pamr_forward_timed": (_int, [_vp, _vp, _vp, _vp, _sz, _int, _int, _int, _int, _int, ctypes.POINTER(_int),
                                      _int, _int, _vp, _vp, _vp]),
    "cl4_peak_extract_scratch_bytes": (_sz, [_int] * 6),
    "cl4_peak_extract": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _int, _int, _int, _int, _int, _int, _vp]),
    "cl4_cam_normalize": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _vp]),
    "cl4_peak_extract_upsampled": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _sz, _int, _int, _int, _int, _int, _int, _vp]),
    "cl4_center_nms_scratch_bytes": (_sz, [_int] * 3),
    "cl4_center_nms": (_int, [_vp, _flt, _flt, _int, _int, _int, _int, _vp, _vp, _int, _vp, _sz, _vp]),
    "cl4_ccl4_scratch_bytes": (_sz, [_int, _int]),
    "cl4_ccl4_components": (_int, [_vp, _vp, _flt, _flt, _flt, _int, _int, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "cl4_smoothing": (_int, [_vp, _vp, _int, _int, _int, _int, _vp]),
    "cl4_pseudo_labels": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _flt, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp,
                                 _int, _int, _int, _int, _vp, _sz, _vp]),
    "cl4_refine_max_contours": (_int, []),
    "cl4_refine_scratch_bytes": (_sz, [_int] * 3),
    "cl4_contours8": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cl4_refine_labels": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, ctypes.c_double, _int, _flt, _int, _int,
                                 ctypes.c_longlong, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _sz, _vp]),
    "cl4_refine_labels_with_point": (_int, [_vp] * 7 + [_int] * 5 + [_vp]),
    "cl4_ins_map_max_instances": (_int, []),
    "cl4_ins_map_scratch_bytes": (_sz, [_int] * 3),
    "cl4_ins_map": (_int, [_vp, _vp, _vp, _vp, _int, _flt, _flt, _flt, _int, _flt, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                           _int, _int, _int, _vp, _sz, _vp]),
    "cl4_ins_masks": (_int, [_vp, _int, _int, _int, _vp, _vp]),
    "cl4_denorm": (_int, [_vp, _vp, _int, _int, ctypes.c_longlong, ctypes.POINTER(_flt), ctypes.POINTER(_flt), _vp]),
    "cl4_denorm_resize_ac": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, ctypes.POINTER(_flt),
                                    ctypes.POINTER(_flt), _vp]),
    "cl4_phase1_scratch_bytes": (_sz, [_int] * 5),
    "cl4_phase1_pseudo_labels": (_int, [_vp, _vp, _vp, ctypes.POINTER(_flt), ctypes.POINTER(_flt), ctypes.POINTER(_int), _int, _int,
                                        _flt, _flt, _flt, _vp, _vp, _vp, _sz, _int, _int, _int, _int, _int, _int, _vp]),
    "cl4_softmax_channels": (_int, [_vp, _vp, _int, _int, ctypes.c_longlong, _vp]),
    "cl4_pseudo_gtmask": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _flt, _flt, _flt, _int, _vp]),
    "cl4_lattice_owner": (_int, [_int, _int, _int, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "cl4_group_pixels": (_int, [_vp, _vp, _int, _int, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
}

_lib = None


def load():
    """Load the CUDA library or fail loudly (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cl4wsis_b200.build` "
                "(nvcc, sm_100a). cl4wsis_b200 has no CPU/PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class Cl4Error(RuntimeError):
    pass


def check(rc, what):
    if rc == CL4_OK:
        return
    msg = load().cl4_last_error().decode("utf-8", "replace")
    if rc == CL4_EINVAL:
        raise ValueError(f"{what}: {msg}")
    if rc == CL4_EUNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise Cl4Error(f"{what} failed (code {rc}): {msg}")


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: cl4wsis_b200 runs on sm_100a only and has no CPU path")


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def float_array(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])
