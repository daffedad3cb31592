"""Drop-in Panoptic-DeepLab post-processing (reference modules/utils.py:463-606; the twin in
dataset/utils.py:623-765 differs only by a ``print``) on hand-written sm_100a kernels.

Same names, argument meaning, return types and error behaviour as the reference:
``find_instance_center``, ``group_pixels``, ``get_instance_segmentation``.  Batched,
sync-free variants used by the pseudo-label pipeline live in ``cl4wsis_b200.pipeline``.
"""
import torch

from .. import _lib


def _heat_plane(ctr_hmp):
    _lib.require_cuda(ctr_hmp, "ctr_hmp")
    if ctr_hmp.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')  # modules/utils.py:476-477
    plane = ctr_hmp.detach().squeeze()
    assert len(plane.size()) == 2, 'Something is wrong with center heatmap dimension.'  # :489
    if plane.dtype != torch.float32:
        plane = plane.float()
    return plane.contiguous()


def _center_nms(plane, threshold, nms_kernel, min_value=0.0):
    """-> (centres int64 [Kc,2] on device, Kc).  One D2H sync to learn Kc, like ``nonzero`` in
    the reference (SURVEY §7.2 'data-dependent K')."""
    lib = _lib.load()
    H, W = plane.shape
    if nms_kernel % 2 == 0:
        # an even kernel makes max_pool2d's output (H-1)x(W-1); the reference then fails at
        # `ctr_hmp[ctr_hmp != ctr_hmp_max_pooled]` (modules/utils.py:485)
        raise RuntimeError(f"find_instance_center: even nms_kernel {nms_kernel} changes the pooled map's shape")
    dev = plane.device
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        nbytes = lib.cl4_center_nms_scratch_bytes(1, H, W)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        cap = min(H * W, 4096)
        while True:
            ctr = torch.empty((cap, 2), dtype=torch.int64, device=dev)
            _lib.check(lib.cl4_center_nms(_lib.ptr(plane), float(threshold), float(min_value), int(nms_kernel), 1, H,
                                          W, _lib.ptr(ctr), _lib.ptr(count), cap, _lib.ptr(scratch), nbytes, st),
                       "find_instance_center")
            n = int(count.item())
            if n <= cap:
                return ctr[:n], n
            cap = n  # rare: more centres than the first guess; run again with room for all of them


def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=5, top_k=None):
    """modules/utils.py:463-502.  ctr_hmp [1,1,H,W] -> LongTensor [Kc,2] of (y,x) in
    ``torch.nonzero`` order.  Does not modify ``ctr_hmp``.

    ``top_k``: as in the reference, all centres are returned when ``top_k`` is None or
    Kc < top_k.  Otherwise the reference's (degenerate) branch is reproduced:
    it prints Kc, takes the ``top_k``-th largest *coordinate* and returns the centres whose
    heat exceeds that number (:498-502).
    """
    plane = _heat_plane(ctr_hmp)
    ctr_all, n = _center_nms(plane, threshold, nms_kernel)
    if top_k is None:
        return ctr_all
    elif n < top_k:
        return ctr_all
    else:
        print(n)
        kth = float(torch.sort(torch.flatten(ctr_all).cpu(), descending=True)[0][top_k - 1])
        ctr, _ = _center_nms(plane, threshold, nms_kernel, min_value=kth)
        return ctr


def group_pixels(ctr, offsets):
    """modules/utils.py:505-542.  ctr LongTensor [Kc,2] (y,x), offsets [1,2,H,W] (dy,dx)
    -> LongTensor [1,H,W] with ids in 1..Kc (first minimum wins ties)."""
    return _group(ctr, offsets, None)


def _group(ctr, offsets, fg):
    lib = _lib.load()
    _lib.require_cuda(offsets, "offsets")
    if offsets.size(0) != 1:
        raise ValueError('Only supports inference for batch size = 1')  # modules/utils.py:516-517
    dev = offsets.device
    off = offsets.detach()
    if off.dtype != torch.float32:
        off = off.float()
    off = off.contiguous()
    if off.dim() != 4 or off.size(1) != 2:
        raise ValueError("offsets must be [1,2,H,W]")
    H, W = off.shape[-2:]
    if not isinstance(ctr, torch.Tensor):
        ctr = torch.as_tensor(ctr)
    ctr = ctr.to(device=dev, dtype=torch.int64).contiguous()
    Kc = ctr.size(0)
    if Kc == 0:
        # torch.argmin over an empty dim-0 raises in the reference (:540)
        raise IndexError("group_pixels: argmin over zero centres")
    fg_u8 = None
    if fg is not None and fg.dtype == torch.bool:  # every reference caller passes a bool mask
        fg_u8 = fg.detach().to(dev).reshape(H, W).to(torch.uint8).contiguous()
    with torch.cuda.device(dev):
        ids = torch.empty((1, H, W), dtype=torch.int64, device=dev)
        _lib.check(lib.cl4_group_pixels(_lib.ptr(ctr), None, Kc, Kc, _lib.ptr(off), _lib.ptr(fg_u8), _lib.ptr(ids),
                                        1, H, W, 0, _lib.stream_ptr(dev)), "group_pixels")
    if fg is not None and fg_u8 is None:  # non-bool fg: literal `(fg * ins_seg).long()` (:606)
        return (fg * ids).long()
    return ids


def get_instance_segmentation(fg, ctr_hmp, offsets, threshold=0.1, nms_kernel=3, top_k=None, ignore=True, beta=5):
    """modules/utils.py:545-606.  fg [1,H,W] bool, ctr_hmp [1,1,H,W], offsets [1,2,H,W]
    -> LongTensor [1,H,W] = fg * instance id.

    beta > 0 adds the centre-clustering merge (:567-594); like the reference it marks merged
    cluster centres with 1.0 in ``ctr_hmp`` IN PLACE and swallows every exception of that
    branch, falling back to the NMS centres (:593-594).
    """
    ctr = find_instance_center(ctr_hmp, threshold=threshold, nms_kernel=nms_kernel, top_k=top_k)

    if beta > 0:  # centre clustering
        try:
            from ..cluster import merge_cluster_centers
            new_ctr = merge_cluster_centers(ctr, ctr_hmp, offsets, fg, beta)
        except Exception:  # noqa: BLE001 - the reference uses a bare except here
            new_ctr = ctr
    else:
        new_ctr = ctr

    if new_ctr.size(0) == 0:  # no peak & no cluster
        if ignore:
            return torch.zeros_like(fg).long()
        else:
            return fg.long()

    return _group(new_ctr, offsets, fg)
