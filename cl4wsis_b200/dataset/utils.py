"""Drop-in validation post-processing (reference dataset/utils.py:623-902) on the sm_100a kernels.

``dataset/utils.py`` carries its own copies of ``find_instance_center`` / ``group_pixels`` /
``get_instance_segmentation`` / ``cluster_peaks`` (:623-793); they differ from the training copies in
``modules/utils.py`` only by the missing ``print`` in the degenerate ``top_k`` branch and by
``MINIMUM_MASK_SIZE = 50`` (:147).  ``get_ins_map`` (:795-902) is what ``Trainer.validate`` calls per
image (train.py:622).
"""
import contextlib
import io

import numpy as np
import torch

from .. import _lib
from ..modules import utils as _mu
from ..modules.utils import group_pixels  # noqa: F401  (identical in both reference files)

MINIMUM_MASK_SIZE = 50   # dataset/utils.py:147
MAXIMUM_NUM_INST = 5     # dataset/utils.py:148


def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=5, top_k=None):
    """dataset/utils.py:623-661 — as modules/utils.py:463-502 without the ``print``."""
    with contextlib.redirect_stdout(io.StringIO()):
        return _mu.find_instance_center(ctr_hmp, threshold=threshold, nms_kernel=nms_kernel, top_k=top_k)


def get_instance_segmentation(fg, ctr_hmp, offsets, threshold=0.1, nms_kernel=3, top_k=None, ignore=True, beta=5):
    """dataset/utils.py:704-765."""
    with contextlib.redirect_stdout(io.StringIO()):
        return _mu.get_instance_segmentation(fg, ctr_hmp, offsets, threshold=threshold, nms_kernel=nms_kernel,
                                             top_k=top_k, ignore=ignore, beta=beta)


def _opencv_label_order(comp, info, n):
    """Order of the contour slots of ONE class as cv2.connectedComponentsWithStats(connectivity=8)
    numbers them: OpenCV's 8-connectivity labelling scans 2x2 blocks in raster order, so labels follow
    (block row of the contour's first pixel, first block column of the contour inside that block row)
    — pinned against cv2 in tests/test_abi_and_host.py.  (Host-side statement of the rule that
    ``cl4_ins_map`` applies on the device; kept for that test.)"""
    H, W = comp.shape
    keys = []
    for s in n:
        y0 = int(info[s, 0]) // W
        br = y0 // 2
        rows = comp[2 * br:2 * br + 2] == s
        keys.append((br, int(np.flatnonzero(rows.any(0))[0]) // 2, s))
    return [k[2] for k in sorted(keys)]


def get_ins_map(out, cls_label, target_size, device, args):
    """post-processing (output -> instance map) — dataset/utils.py:795-902, on the device.

    out: dict with 'seg' [B,C+1,H,W] logits, 'center' [B,C,H,W], 'offset' [B,2,H,W] (B = 2 with
    ``args.val_flip``); returns (seg_map [H,W] int64 ndarray, pred_label [n], pred_mask [n,H,W] bool,
    pred_score [n] float64), instances ordered as the reference's loops produce them.  Like the reference
    it rescales ``out['offset'][0]`` IN PLACE (:831-832).  One ``cl4_ins_map`` call does the softmax, the
    flip averaging, the label cleaning, the argmax, the contours of every class, per-contour centre NMS,
    clustering, grouping and the per-instance scores; the host reads the instance count once and a second
    launch writes the boolean masks.  (The reference's empty-result branch uses ``np.bool``, which newer
    numpy removed; ``np.bool_`` is used here.)

    Capacity (a clear error instead of a silently different answer): 1024 contours of >= 50 px, 4096 NMS
    centres / cluster blobs / instances per image, 64 accepted cluster centres per contour; the number of NMS
    centres inside one contour is not limited.
    """
    lib = _lib.load()
    seg, ctr, off = out['seg'].detach(), out['center'].detach(), out['offset']
    for name, t in (("out['seg']", seg), ("out['center']", ctr), ("out['offset']", off)):
        _lib.require_cuda(t, name)
        if t.dtype != torch.float32:
            raise TypeError(f"get_ins_map: {name} must be float32")
    flip = bool(args.val_flip)
    if seg.shape[0] < (2 if flip else 1):
        raise IndexError("get_ins_map: val_flip needs the mirrored view as out[...][1]")
    seg, ctr = seg.contiguous(), ctr.contiguous()
    off0 = off.detach()[0]
    if not off0.is_contiguous():
        raise ValueError("get_ins_map: out['offset'][0] must be contiguous (it is rescaled in place)")
    C, (H, W) = ctr.shape[1], seg.shape[-2:]
    dev = seg.device
    lab = None
    if args.val_clean:
        lab = cls_label[0].detach().to(device=dev, dtype=torch.float32).contiguous()
    cap = lib.cl4_ins_map_max_instances()
    with torch.cuda.device(dev):
        seg_map = torch.empty((H, W), dtype=torch.int64, device=dev)
        inst_map = torch.empty((H, W), dtype=torch.int32, device=dev)
        labels = torch.empty(cap, dtype=torch.int32, device=dev)
        scores = torch.empty(cap, dtype=torch.float64, device=dev)
        head = torch.zeros(2, dtype=torch.int32, device=dev)          # [n, status]
        nbytes = lib.cl4_ins_map_scratch_bytes(C, H, W)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        st = _lib.stream_ptr(dev)
        _lib.check(lib.cl4_ins_map(_lib.ptr(seg), _lib.ptr(ctr), _lib.ptr(off0), _lib.ptr(lab), int(flip),
                                   float(target_size[0] / H), float(target_size[1] / W), float(args.val_thresh),
                                   int(args.val_kernel), float(args.beta), 1 if args.val_ignore else 0, MINIMUM_MASK_SIZE,
                                   _lib.ptr(seg_map), _lib.ptr(inst_map), _lib.ptr(labels), _lib.ptr(scores),
                                   _lib.ptr(head), _lib.ptr(head[1:]), C, H, W, _lib.ptr(scratch), nbytes, st), "get_ins_map")
        n, status = head.tolist()                                       # the one host synchronisation
        if status:
            raise NotImplementedError(f"get_ins_map: capacity exceeded (status {status}: 1 = more than 1024 contours, "
                                      f"2 = more than 64 cluster centres in a contour, 4 = more than {cap} centres or instances)")
        if n == 0:
            return (seg_map.cpu().numpy(), np.stack([0], 0), np.stack([np.zeros(tuple(target_size), dtype=np.bool_)], 0),
                    np.stack([0], 0))
        masks = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
        _lib.check(lib.cl4_ins_masks(_lib.ptr(inst_map), n, H, W, _lib.ptr(masks), st), "get_ins_map masks")
    return (seg_map.cpu().numpy(), labels[:n].cpu().numpy().astype(np.int64), masks.cpu().numpy().astype(np.bool_),
            scores[:n].cpu().numpy())
