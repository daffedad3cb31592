"""ctypes front-end of ``cl4_oracle.c`` — numpy in, numpy out.

TEST INFRASTRUCTURE ONLY (see ``cl4_oracle.c`` header).  Parity status: pinned
against reference-generated fixtures in ``tests/golden`` (``make_golden.py``).

Function names and argument meaning follow the reference:
  PAMR.forward                -> pamr()                      wss/modules.py:133-152
  peak_extract                -> peak_extract()              wss/utils.py:3-25
  find_instance_center        -> find_instance_center()      modules/utils.py:463-502
  group_pixels                -> group_pixels()              modules/utils.py:505-542
  get_instance_segmentation   -> get_instance_segmentation() modules/utils.py:545-606 (beta<=0 path)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcl4_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_longlong)
_u8p = ctypes.POINTER(ctypes.c_ubyte)


def build(force=False):
    """Compile ``libcl4_oracle.so`` with the committed Makefile (gcc + OpenMP)."""
    src = os.path.join(_HERE, "cl4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "libcl4_oracle.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        lib.cl4o_num_threads.restype = ctypes.c_int
        lib.cl4o_set_num_threads.argtypes = [ctypes.c_int]
        lib.cl4o_resize_bilinear_ac.argtypes = [_f32p, _f32p] + [ctypes.c_int] * 5
        lib.cl4o_local_affinity.argtypes = [_f32p, _f32p] + [ctypes.c_int] * 3 + [_i32p, ctypes.c_int, ctypes.c_int]
        lib.cl4o_local_stdev.argtypes = [_f32p, _f32p] + [ctypes.c_int] * 3 + [_i32p, ctypes.c_int]
        lib.cl4o_pamr_weights.argtypes = [_f32p, _f32p] + [ctypes.c_int] * 4 + [_i32p, ctypes.c_int]
        lib.cl4o_pamr.argtypes = [_f32p, _f32p, _f32p] + [ctypes.c_int] * 7 + [_i32p, ctypes.c_int, ctypes.c_int]
        lib.cl4o_peak_extract.argtypes = [_f32p, _f32p, _i32p, _i32p] + [ctypes.c_int] * 6
        lib.cl4o_center_nms.argtypes = [_f32p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        _i64p, ctypes.c_longlong]
        lib.cl4o_center_nms.restype = ctypes.c_longlong
        lib.cl4o_group_pixels.argtypes = [_i64p, ctypes.c_int, _f32p, _u8p, _i64p, ctypes.c_int, ctypes.c_int]
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _check(rc, what):
    if rc < 0:
        raise RuntimeError(f"oracle {what} failed with code {rc}")


def num_threads():
    return int(_load().cl4o_num_threads())


def set_num_threads(n):
    _load().cl4o_set_num_threads(int(n))


def resize_bilinear_ac(x, size):
    x = _f32(x)
    h, w = x.shape[-2:]
    H, W = size
    out = np.empty(x.shape[:-2] + (H, W), np.float32)
    planes = int(np.prod(x.shape[:-2], dtype=np.int64))
    _check(_load().cl4o_resize_bilinear_ac(_ptr(x, _f32p), _ptr(out, _f32p), planes, h, w, H, W), "resize")
    return out


def local_affinity(x, dilations, mode=0):
    """LocalAffinity (mode 0), LocalAffinityAbs (1), LocalAffinityCopy (2) .forward(x):
    wss/modules.py:47-62, :115-119, :65-83.  x [B,K,H,W] -> [B,K,8*D,H,W]."""
    x = _f32(x)
    B, K, H, W = x.shape
    dil = np.ascontiguousarray(dilations, dtype=np.int32)
    out = np.empty((B, K, 8 * len(dil), H, W), np.float32)
    _check(_load().cl4o_local_affinity(_ptr(x, _f32p), _ptr(out, _f32p), B * K, H, W, _ptr(dil, _i32p), len(dil),
                                       int(mode)), "local_affinity")
    return out


def local_stdev(x, dilations):
    """LocalStDev.forward(x): wss/modules.py:86-112.  x [B,K,H,W] -> [B,K,1,H,W]."""
    x = _f32(x)
    B, K, H, W = x.shape
    dil = np.ascontiguousarray(dilations, dtype=np.int32)
    out = np.empty((B, K, 1, H, W), np.float32)
    _check(_load().cl4o_local_stdev(_ptr(x, _f32p), _ptr(out, _f32p), B * K, H, W, _ptr(dil, _i32p), len(dil)),
           "local_stdev")
    return out


def pamr_weights(x, dilations):
    x = _f32(x)
    B, K, H, W = x.shape
    dil = np.ascontiguousarray(dilations, dtype=np.int32)
    w = np.empty((B, 8 * len(dil), H, W), np.float32)
    _check(_load().cl4o_pamr_weights(_ptr(x, _f32p), _ptr(w, _f32p), B, K, H, W, _ptr(dil, _i32p), len(dil)),
           "pamr_weights")
    return w


def pamr(x, mask, num_iter=10, dilations=(1, 2, 4, 8, 12, 24)):
    """PAMR(num_iter, dilations).forward(x, mask) -> [B,C,H,W] float32."""
    x, mask = _f32(x), _f32(mask)
    B, K, H, W = x.shape
    Bm, C, h, w = mask.shape
    if Bm != B:
        raise ValueError("batch mismatch")
    dil = np.ascontiguousarray(dilations, dtype=np.int32)
    out = np.empty((B, C, H, W), np.float32)
    _check(_load().cl4o_pamr(_ptr(x, _f32p), _ptr(mask, _f32p), _ptr(out, _f32p), B, K, C, H, W, h, w,
                             _ptr(dil, _i32p), len(dil), int(num_iter)), "pamr")
    return out


def peak_extract(heat, kernel=5, K=25):
    """-> (scores f32 [B,C,K], ys i32, xs i32), sorted by (score desc, flat index asc)."""
    heat = _f32(heat)
    B, C, H, W = heat.shape
    sc = np.empty((B, C, K), np.float32)
    ys = np.empty((B, C, K), np.int32)
    xs = np.empty((B, C, K), np.int32)
    _check(_load().cl4o_peak_extract(_ptr(heat, _f32p), _ptr(sc, _f32p), _ptr(ys, _i32p), _ptr(xs, _i32p),
                                     B, C, H, W, int(kernel), int(K)), "peak_extract")
    return sc, ys, xs


def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=5, top_k=None):
    """-> int64 [Kc,2] (y,x) in row-major order.  ``top_k`` follows the reference,
    including its degenerate branch: when Kc >= top_k the reference
    thresholds the heat-map by a *coordinate* value (modules/utils.py:500-502)."""
    ctr_hmp = _f32(ctr_hmp)
    if ctr_hmp.shape[0] != 1:
        raise ValueError("Only supports inference for batch size = 1")
    plane = np.squeeze(ctr_hmp)
    assert plane.ndim == 2, "Something is wrong with center heatmap dimension."
    H, W = plane.shape
    plane = np.ascontiguousarray(plane)
    cap = H * W
    ctr = np.empty((cap, 2), np.int64)
    n = _load().cl4o_center_nms(_ptr(plane, _f32p), float(threshold), int(nms_kernel), H, W, _ptr(ctr, _i64p), cap)
    _check(n, "center_nms")
    ctr = ctr[:n].copy()
    if top_k is None or n < top_k:
        return ctr
    # degenerate top-k branch: the k-th largest *coordinate* becomes a heat threshold
    kth = np.sort(ctr.reshape(-1))[::-1][top_k - 1]
    nms = _nms_plane(plane, threshold, nms_kernel)
    return np.argwhere(nms > np.float32(kth)).astype(np.int64)


def _nms_plane(plane, threshold, nms_kernel):
    """thresholded + suppressed heat-map (values -1 where not a centre)."""
    H, W = plane.shape
    ctr = np.empty((H * W, 2), np.int64)
    n = _load().cl4o_center_nms(_ptr(plane, _f32p), float(threshold), int(nms_kernel), H, W, _ptr(ctr, _i64p), H * W)
    out = np.full((H, W), -1.0, np.float32)
    c = ctr[:n]
    out[c[:, 0], c[:, 1]] = plane[c[:, 0], c[:, 1]]
    return out


def group_pixels(ctr, offsets, fg=None):
    """-> int64 [1,H,W], ids in 1..Kc (times fg when given)."""
    offsets = _f32(offsets)
    if offsets.shape[0] != 1:
        raise ValueError("Only supports inference for batch size = 1")
    ctr = np.ascontiguousarray(ctr, dtype=np.int64)
    _, _, H, W = offsets.shape
    ids = np.empty((1, H, W), np.int64)
    fgp = None
    if fg is not None:
        fg = np.ascontiguousarray(np.asarray(fg).reshape(H, W) != 0, dtype=np.uint8)
        fgp = _ptr(fg, _u8p)
    _check(_load().cl4o_group_pixels(_ptr(ctr, _i64p), ctr.shape[0], _ptr(offsets, _f32p), fgp,
                                     _ptr(ids, _i64p), H, W), "group_pixels")
    return ids


def cluster_peaks(offset_map, fg, thresh=2.5, beta=5):
    """modules/utils.py:608-632: 4-connected components (OpenCV, as the reference) of the weak-offset
    foreground; centroids (y,x) of the components with 21-beta < area < 21+beta, label 0 included."""
    import cv2
    off = _f32(offset_map)
    mag = np.sqrt(off[1] * off[1] + off[0] * off[0])
    weak = ((mag < thresh) & (np.asarray(fg) != 0)).astype(np.uint8)
    n, _, stats, cent = cv2.connectedComponentsWithStats(weak, connectivity=4)
    return np.int32([cent[k][::-1] for k in range(n) if 21 - beta < stats[k, cv2.CC_STAT_AREA] < 21 + beta])


def get_instance_segmentation(fg, ctr_hmp, offsets, threshold=0.1, nms_kernel=3, top_k=None, ignore=True, beta=0):
    """modules/utils.py:545-606.  Returns (ids, heat_after): the reference marks accepted cluster
    centres with 1.0 in ctr_hmp in place; the oracle returns the marked copy instead."""
    ctr_hmp = _f32(ctr_hmp).copy()
    offsets = _f32(offsets)
    fg = np.asarray(fg)
    ctr = find_instance_center(ctr_hmp, threshold, nms_kernel, top_k)
    new_ctr = ctr
    if beta > 0:
        cl = cluster_peaks(offsets[0], fg[0], beta=beta)
        cl = np.array([[cy, cx] for cy, cx in cl if ctr_hmp[0, 0, cy, cx] > 0.05], dtype=np.int64).reshape(-1, 2)
        if len(cl):
            if len(ctr) == 0:
                new_ctr = cl
                ctr_hmp[0, 0, cl[:, 0], cl[:, 1]] = 1.0
            else:
                keep = []
                for c in cl:
                    d = np.sqrt(((ctr.astype(np.float32) - c.astype(np.float32)) ** 2).sum(-1, dtype=np.float32)).min()
                    if d > 100:
                        keep.append(c)
                        ctr_hmp[0, 0, c[0], c[1]] = 1.0
                if keep:
                    new_ctr = np.concatenate([ctr, np.array(keep, np.int64)], 0)
    if new_ctr.shape[0] == 0:
        ids = np.zeros(fg.shape, np.int64) if ignore else fg.astype(np.int64)
    else:
        ids = group_pixels(new_ctr, offsets, fg=fg)
    return ids, ctr_hmp
