"""Centre clustering for ``get_instance_segmentation`` (reference modules/utils.py:567-594 and
``cluster_peaks`` :608-632) with the connected components computed on the GPU.

The reference copies offsets and fg to the host and runs OpenCV there; here ``cl4_ccl4_components``
labels the weak-offset foreground on the device and returns the area / coordinate sums of the
components whose area passes the ``21 - beta < area < 21 + beta`` filter, in OpenCV's label order.
Only those few numbers cross to the host, where the (tiny) merge rules run.
"""
import numpy as np
import torch

from . import _lib


def cluster_peaks(offsets, fg, thresh=2.5, beta=5):
    """offsets [1,2,H,W] (dy,dx) fp32 CUDA, fg [1,H,W] bool CUDA -> int32 ndarray [n,2] of (y,x)
    centroids, truncated like ``np.int32(centroids)`` — modules/utils.py:608-632."""
    lib = _lib.load()
    _lib.require_cuda(offsets, "offsets")
    off = offsets.detach()[0]
    if off.dtype != torch.float32:
        off = off.float()
    off = off.contiguous()
    H, W = off.shape[-2:]
    dev = off.device
    fg_u8 = (fg.detach().to(dev).reshape(H, W) != 0).to(torch.uint8).contiguous()
    with torch.cuda.device(dev):
        nbytes = lib.cl4_ccl4_scratch_bytes(H, W)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        cap = 1024
        while True:
            roots = torch.empty((cap, 2), dtype=torch.int64, device=dev)
            stats = torch.empty((cap + 1, 3), dtype=torch.int64, device=dev)
            _lib.check(lib.cl4_ccl4_components(_lib.ptr(off), _lib.ptr(fg_u8), float(thresh), float(21 - beta),
                                               float(21 + beta), H, W, _lib.ptr(roots), _lib.ptr(stats),
                                               _lib.ptr(count), cap, _lib.ptr(scratch), nbytes,
                                               _lib.stream_ptr(dev)), "cluster_peaks")
            n = int(count.item())
            if n <= cap:
                break
            cap = n
    st = stats[: n + 1].cpu().numpy()
    peaks = []
    # OpenCV's label 0 is everything that is not a component; the reference filters it like any other
    # label (k starts at 0, modules/utils.py:630)
    for k in range(n + 1):
        area, sx, sy = int(st[k, 0]), int(st[k, 1]), int(st[k, 2])
        if k == 0 and not (21 - beta < area < 21 + beta):
            continue
        if area == 0:
            continue
        peaks.append([float(sy) / float(area), float(sx) / float(area)])  # centroid in double, (y, x)
    return np.int32(peaks)


def merge_cluster_centers(ctr, ctr_hmp, offsets, fg, beta):
    """The merge of NMS centres and cluster centres, modules/utils.py:569-592.  Marks accepted
    cluster centres with 1.0 in ``ctr_hmp`` in place, as the reference does."""
    cand = cluster_peaks(offsets, fg, beta=beta)
    if len(cand):
        pts = torch.from_numpy(cand.astype(np.int64)).to(ctr_hmp.device)
        heat = ctr_hmp[0, 0][pts[:, 0], pts[:, 1]].float().cpu().numpy()
        cand = cand[heat > 0.05]  # only cluster centres the heat-map supports (:571)
    if len(cand) == 0:
        return ctr.clone()
    if ctr.size(0) == 0:
        accepted = cand  # no NMS centre at all: the cluster centres are the centres (:578-583)
    else:
        # keep a cluster centre when it is farther than 100 px from EVERY NMS centre (:585-591);
        # fp32 like ``torch.norm(ctr.float() - c.float(), dim=-1).min()``
        nms = ctr.cpu().numpy().astype(np.float32)
        diff = nms[None, :, :] - cand[:, None, :].astype(np.float32)
        dmin = np.sqrt((diff * diff).sum(-1, dtype=np.float32)).min(axis=1)
        accepted = cand[dmin > 100]
    if len(accepted) == 0:
        return ctr.clone()
    acc = torch.from_numpy(accepted.astype(np.int64)).to(ctr.device)
    ctr_hmp[0, 0][acc[:, 0].to(ctr_hmp.device), acc[:, 1].to(ctr_hmp.device)] = 1.0  # mark as new peak
    return acc if ctr.size(0) == 0 else torch.cat([ctr, acc], dim=0)
