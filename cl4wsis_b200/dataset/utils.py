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
    — pinned against cv2 in tests/test_abi_and_host.py."""
    H, W = comp.shape
    keys = []
    for s in n:
        y0 = int(info[s, 0]) // W
        br = y0 // 2
        rows = comp[2 * br:2 * br + 2] == s
        keys.append((br, int(np.flatnonzero(rows.any(0))[0]) // 2, s))
    return [k[2] for k in sorted(keys)]


def get_ins_map(out, cls_label, target_size, device, args):
    """post-processing (output -> instance map) — dataset/utils.py:795-902.

    out: dict with 'seg' [B,C+1,H,W] logits, 'center' [B,C,H,W], 'offset' [B,2,H,W] (B = 2 with
    ``args.val_flip``); returns (seg_map [H,W] int64 ndarray, pred_label [n], pred_mask [n,H,W] bool,
    pred_score [n]).  Like the reference it rescales ``out['offset'][0]`` IN PLACE (:826-828).
    Contours come from the GPU (``cl4_contours8``) in OpenCV's label order, so the instance lists
    are ordered as the reference's.  (The reference's empty-result branch uses ``np.bool``, which
    newer numpy removed; ``np.bool_`` is used here.)
    """
    pred_label, pred_mask, pred_score = [], [], []

    seg_prob = torch.softmax(out['seg'].detach(), 1)
    center_map = out['center'].detach()
    offset_map = out['offset'][0].detach()

    if args.val_flip:
        seg_prob = (seg_prob[0] + seg_prob[1].flip(-1)) / 2.
        center_map = (center_map[0] + center_map[1].flip(-1)) / 2.
    else:
        seg_prob = seg_prob[0]
        center_map = center_map[0]

    out_size = seg_prob.shape[1:]
    offset_map[0, :, :] = offset_map[0, :, :] * (target_size[0] / out_size[0])
    offset_map[1, :, :] = offset_map[1, :, :] * (target_size[1] / out_size[1])

    if args.val_clean:
        seg_prob[1:, :, :] *= cls_label[0, :, None, None].to(device)

    seg_map = torch.argmax(seg_prob, 0)
    C = center_map.shape[0]

    # 8-connected contours of every class present, on the device (the reference: one cv2 call per class)
    comp, info, ncomp = _mu.contours8(seg_map[None], torch.ones((1, C), device=seg_map.device), min_area=MINIMUM_MASK_SIZE)
    n = int(ncomp[0])
    comp_h, info_h = comp[0].cpu().numpy(), info[0, :n].cpu().numpy()
    center_map = center_map.float()
    offset_f = offset_map.float()

    for cls in sorted(set(int(c) for c in info_h[:, 1])):  # torch.unique(seg_map) - 1, ascending
        slots = [s for s in range(n) if int(info_h[s, 1]) == cls]
        for s in _opencv_label_order(comp_h, info_h, slots):
            contour_mask = comp[0] == s
            center_map_cls_roi = center_map[cls] * contour_mask
            ins_map = get_instance_segmentation(contour_mask[None, ...], center_map_cls_roi[None, None, ...],
                                                offset_f[None, ...], threshold=args.val_thresh,
                                                nms_kernel=args.val_kernel, beta=args.beta, ignore=args.val_ignore)
            ins_map = ins_map.squeeze(0)
            n_ins = int(ins_map.max())
            for id in range(1, n_ins + 1):
                mask = (ins_map == id)
                if mask.sum() > 0:
                    index = torch.where(mask)
                    center_idx = center_map_cls_roi[index].argmax()
                    seg_score = seg_prob[cls + 1][index].mean().item()
                    cy, cx = index[0][center_idx], index[1][center_idx]
                    center_score = center_map_cls_roi[cy, cx].item()
                    if center_score >= 1:  # clustered centre: conf = seg_score
                        center_score = seg_score
                    pred_label.append(cls)
                    pred_mask.append(mask.cpu().numpy())
                    pred_score.append(center_score * seg_score)

    if len(pred_label) == 0:
        pred_label.append(0)
        pred_mask.append(np.zeros(target_size, dtype=np.bool_))
        pred_score.append(0)

    return seg_map.cpu().numpy(), np.stack(pred_label, 0), np.stack(pred_mask, 0), np.stack(pred_score, 0)
