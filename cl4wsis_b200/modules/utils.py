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


# ------------------------------------------------------------------------------------------------
# refine_label_generation — modules/utils.py:257-385 (the phase-2 caller, train.py:492-500)
# ------------------------------------------------------------------------------------------------
MINIMUM_MASK_SIZE = 20   # modules/utils.py:14 (the validation twin dataset/utils.py:147 uses 50)
MAXIMUM_NUM_INST = 5     # modules/utils.py:15

_GAUSS_CACHE = {}


def gaussian(sigma=6):
    """The (6*sigma+3)^2 float64 bump of modules/utils.py:49-59 (numpy, as in the reference)."""
    import numpy as np
    size = 6 * sigma + 3
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0, y0 = 3 * sigma + 1, 3 * sigma + 1
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))


def _gauss_f32(sigma, device):
    key = (sigma, str(device))
    if key not in _GAUSS_CACHE:
        # the reference max-combines the float64 bump into a float32 map: same as using float32(bump)
        _GAUSS_CACHE[key] = torch.from_numpy(gaussian(sigma).astype("float32")).to(device).contiguous()
    return _GAUSS_CACHE[key]


def _refine_inputs(seg_map, center_map, offset_map, label, gt_seg_map):
    for t, name in ((seg_map, "seg_map"), (center_map, "center_map"), (offset_map, "offset_map"),
                    (gt_seg_map, "gt_seg_map")):
        _lib.require_cuda(t, name)
    dev = center_map.device
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    return f32(seg_map), f32(center_map), f32(offset_map), f32(label), gt_seg_map.detach().to(dev).long().contiguous()


def refine_label_generation_device(seg_map, center_map, offset_map, label, gt_seg_map, top_k, args):
    """The whole batch in one sequence of kernels (``cl4_refine_labels``), no host round trip.
    Returns (dict(center, offset, weight), status) where ``status`` is a device int32 tensor: non-zero
    means a capacity limit was hit and the result must be recomputed with the per-contour path."""
    lib = _lib.load()
    seg, ctr, off, lab, gt = _refine_inputs(seg_map, center_map, offset_map, label, gt_seg_map)
    B, C, H, W = ctr.shape
    if seg.shape != (B, C + 1, H, W) or off.shape != (B, 2, H, W) or lab.shape != (B, C) or gt.shape != (B, H, W):
        raise ValueError("refine_label_generation: seg [B,C+1,H,W], center [B,C,H,W], offset [B,2,H,W], "
                         "label [B,C], gt_seg [B,H,W] expected")
    dev = ctr.device
    sigma = int(args.sigma)
    if sigma != args.sigma:
        raise NotImplementedError("refine_label_generation: integer sigma only (argparser.py:221)")
    with torch.cuda.device(dev):
        out_c = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        out_o = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
        out_w = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if B:
            nbytes = lib.cl4_refine_scratch_bytes(B, H, W)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            g = _gauss_f32(sigma, dev)
            _lib.check(lib.cl4_refine_labels(
                _lib.ptr(seg), _lib.ptr(ctr), _lib.ptr(off), _lib.ptr(lab), _lib.ptr(gt), _lib.ptr(g), sigma,
                float(args.refine_thresh), int(args.kernel), float(args.beta), MINIMUM_MASK_SIZE, MAXIMUM_NUM_INST,
                -1 if top_k is None else int(top_k), _lib.ptr(out_c), _lib.ptr(out_o), _lib.ptr(out_w),
                _lib.ptr(status), B, C, H, W, _lib.ptr(scratch), nbytes, _lib.stream_ptr(dev)), "refine_label_generation")
    return {'center': out_c, 'offset': out_o, 'weight': out_w}, status


def contours8(gt_seg_map, label, min_area=MINIMUM_MASK_SIZE):
    """8-connected contours of every valid (image, class) on the GPU (``cl4_contours8``):
    -> (comp [B,H,W] int32 slot map, info [B,n_max,5] int32 (first pixel, cls, cx, cy, area), ncomp [B]).
    Hard limit: ``cl4_refine_max_contours()`` (1024) contours of >= ``min_area`` px per image, ``NotImplementedError`` beyond."""
    lib = _lib.load()
    _lib.require_cuda(gt_seg_map, "gt_seg_map")
    dev = gt_seg_map.device
    gt = gt_seg_map.detach().long().contiguous()
    lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, H, W = gt.shape
    C = lab.shape[1]
    nmax = lib.cl4_refine_max_contours()
    with torch.cuda.device(dev):
        comp = torch.empty((B, H, W), dtype=torch.int32, device=dev)
        info = torch.zeros((B, nmax, 5), dtype=torch.int32, device=dev)
        ncomp = torch.zeros(B, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if B:
            nbytes = lib.cl4_refine_scratch_bytes(B, H, W)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.cl4_contours8(_lib.ptr(gt), _lib.ptr(lab), int(min_area), B, C, H, W, _lib.ptr(comp),
                                         _lib.ptr(info), _lib.ptr(ncomp), _lib.ptr(status), _lib.ptr(scratch), nbytes,
                                         _lib.stream_ptr(dev)), "contours8")
    if int(status.item()) != 0:
        raise NotImplementedError(f"contours8: more than {nmax} contours of >= {min_area} px in one image")
    return comp, info, ncomp


def refine_label_generation_per_contour(seg_map, center_map, offset_map, label, gt_seg_map, top_k, args):
    """The reference's loop structure (image x contour x instance, modules/utils.py:295-377) on this
    package's kernels, at the reference's cost of several host round trips per contour.  Used when the batched
    path reports a capacity overflow.  Exact for every input with at most ``cl4_refine_max_contours()`` (1024)
    contours of >= 20 px per image; beyond that ``contours8`` raises ``NotImplementedError`` (OpenCV, which the
    reference uses, has no such limit -- a 512 x 512 label map would need a contour every 256 px to get there)."""
    seg, ctr, off, lab, gt = _refine_inputs(seg_map, center_map, offset_map, label, gt_seg_map)
    B, C, H, W = ctr.shape
    dev = ctr.device
    prob = torch.softmax(seg, dim=1)
    prob[:, 1:] *= lab[:, :, None, None]
    out_c = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev)
    out_o = torch.zeros((B, 2, H, W), dtype=torch.float32, device=dev)
    out_w = torch.zeros((B, 1, H, W), dtype=torch.float32, device=dev)
    yy = torch.arange(H, dtype=torch.float32, device=dev).view(H, 1).expand(H, W)
    xx = torch.arange(W, dtype=torch.float32, device=dev).view(1, W).expand(H, W)
    sigma = args.sigma
    g = _gauss_f32(int(sigma), dev)
    comp, info, ncomp = contours8(gt, lab)
    info_h, ncomp_h = info.cpu().numpy(), ncomp.cpu().numpy()
    for b in range(B):
        for s in range(int(ncomp_h[b])):
            _, cls, cx, cy, _ = (int(v) for v in info_h[b, s])
            contour = comp[b] == s
            heat = ctr[b, cls] * contour
            ins = get_instance_segmentation(contour[None], heat[None, None], off[b][None],
                                            threshold=args.refine_thresh, nms_kernel=args.kernel, ignore=True,
                                            beta=args.beta, top_k=top_k).squeeze(0)
            n_ins = int(ins.max())
            if n_ins > MAXIMUM_NUM_INST:
                continue
            for i in range(1, n_ins + 1):
                index = torch.where(ins == i)
                if index[0].numel() == 0:
                    continue
                pmax = heat[index].argmax()
                seg_score = prob[b, cls + 1][index].mean().item()
                py, px = index[0][pmax].item(), index[1][pmax].item()
                center_score = heat[py, px].item()
                if center_score < args.refine_thresh:
                    py, px = cy, cx
                    conf = seg_score
                else:
                    conf = center_score * seg_score
                conf = max(0, min(conf, 1))
                # gaussian max-splat (center_map_gen, modules/utils.py:84-119)
                x0, y0 = int(round(px - 3 * sigma - 1)), int(round(py - 3 * sigma - 1))
                x1, y1 = int(round(px + 3 * sigma + 2)), int(round(py + 3 * sigma + 2))
                ix0, ix1, iy0, iy1 = max(0, x0), min(x1, W), max(0, y0), min(y1, H)
                out_c[b, cls, iy0:iy1, ix0:ix1] = torch.maximum(out_c[b, cls, iy0:iy1, ix0:ix1],
                                                                g[iy0 - y0:iy1 - y0, ix0 - x0:ix1 - x0])
                out_w[b, 0][index] = conf
                out_o[b, 0][index] = py - yy[index]
                out_o[b, 1][index] = px - xx[index]
    return {'center': out_c, 'offset': out_o, 'weight': out_w}


def refine_label_generation(seg_map, center_map, offset_map, label, gt_seg_map, top_k, args):
    """Refined-label generation (Self-Refinement) with image-level labels — modules/utils.py:257-385.

    seg_map [B,C+1,H,W] logits, center_map [B,C,H,W], offset_map [B,2,H,W], label [B,C] one-hot,
    gt_seg_map [B,H,W] labels, ``args`` with ``refine_thresh``, ``kernel``, ``beta``, ``sigma``
    -> {'center': [B,C,H,W], 'offset': [B,2,H,W], 'weight': [B,1,H,W]} float32 on the inputs' device.

    Runs the whole batch on the device with ONE host synchronisation (the status word); the
    reference needs >= 6 per contour.  Inputs that exceed a capacity limit of the
    batched path (status != 0) are recomputed by the per-contour path (which shares the hard limit of 1024
    contours of >= 20 px per image and raises ``NotImplementedError`` beyond it).
    """
    out, status = refine_label_generation_device(seg_map, center_map, offset_map, label, gt_seg_map, top_k, args)
    if int(status.item()) != 0:
        return refine_label_generation_per_contour(seg_map, center_map, offset_map, label, gt_seg_map, top_k, args)
    return out


def refine_label_generation_with_point(seg_map, gt_point_cls, offset_map, label, gt_seg_map, args=None):
    """Refined-label generation with point labels — modules/utils.py:388-460, same signature and result dict
    ({'offset': [B,2,H,W], 'weight': [B,1,H,W]} float32 on the inputs' device).

    gt_point_cls [B,C,MAX_NUM_POINTS,2] (y, x), rows with y == 0 or x == 0 are padding (the reference's filter, :436);
    ``seg_map`` is only shape-checked (the reference computes a softmax of it that it never uses, :412-413) and ``args`` is
    unused there as well.  One launch for the whole batch, no host synchronisation (the reference: a `group_pixels` pass per
    (image, class) and a `torch.where` per instance)."""
    lib = _lib.load()
    _lib.require_cuda(offset_map, "offset_map")
    dev = offset_map.device
    off = offset_map.detach()
    if off.dtype != torch.float32:
        raise TypeError("refine_label_generation_with_point: offset_map must be float32")
    off = off.contiguous()
    B, two, H, W = off.shape
    if two != 2 or seg_map.shape[0] != B or tuple(seg_map.shape[-2:]) != (H, W):
        raise ValueError("refine_label_generation_with_point: seg_map [B,C+1,H,W] and offset_map [B,2,H,W] do not match")
    lab = label.detach().to(device=dev, dtype=torch.float32).contiguous()
    C = lab.shape[1]
    gt = gt_seg_map.detach().to(device=dev, dtype=torch.int64).contiguous()
    p = gt_point_cls.detach().to(dev)
    if p.dim() != 4 or p.shape[0] != B or p.shape[1] != C or p.shape[3] != 2:
        raise ValueError("refine_label_generation_with_point: gt_point_cls must be [B,C,MAX_NUM_POINTS,2]")
    M = p.shape[2]
    keep = ((p[..., 0] != 0) & (p[..., 1] != 0)).to(torch.uint8).contiguous()   # the filter of :436 on the caller's values
    pts = p.to(torch.int32).to(torch.int64).contiguous()                          # np.int32(...) then .long() (:437, :442)
    with torch.cuda.device(dev):
        out_o = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
        out_w = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        _lib.check(lib.cl4_refine_labels_with_point(_lib.ptr(gt), _lib.ptr(lab), _lib.ptr(pts), _lib.ptr(keep), _lib.ptr(off),
                                                    _lib.ptr(out_o), _lib.ptr(out_w), B, C, M, H, W, _lib.stream_ptr(dev)),
                   "refine_label_generation_with_point")
    return {'offset': out_o, 'weight': out_w}


def pseudo_label_generation_batch(seg_gt, peaks, cls_label, pseudo_thresh, sigma):
    """The per-image loop of train.py:451-477 — points from ``peak_extract`` filtered by
    ``pseudo_thresh``, then ``pseudo_label_generation`` (modules/utils.py:179-253) per image — for the
    whole batch on the device.

    seg_gt [B,H,W] labels (CUDA), peaks = (conf f32, ys i32, xs i32) each [B,C,K] CUDA tensors as
    returned by ``cl4wsis_b200.wss.utils.peak_extract_device``, cls_label [B,C] (non-zero = class
    considered, train.py:446-447) -> (center [B,C,H,W], offset [B,2,H,W], weight [B,1,H,W],
    total_match [B] int32), float32 on the device: the stacked outputs of the reference loop.
    """
    lib = _lib.load()
    _lib.require_cuda(seg_gt, "seg_gt")
    conf, ys, xs = peaks
    dev = seg_gt.device
    gt = seg_gt.detach().long().contiguous()
    lab = torch.as_tensor(cls_label).detach().to(device=dev, dtype=torch.float32).contiguous()
    conf = conf.detach().to(device=dev, dtype=torch.float32).contiguous()
    ys = ys.detach().to(device=dev, dtype=torch.int32).contiguous()
    xs = xs.detach().to(device=dev, dtype=torch.int32).contiguous()
    B, H, W = gt.shape
    C = lab.shape[1]
    if conf.shape[:2] != (B, C) or ys.shape != conf.shape or xs.shape != conf.shape:
        raise ValueError("pseudo_label_generation_batch: peaks must be three [B,C,K] tensors")
    K = conf.shape[2]
    sigma = int(sigma)
    with torch.cuda.device(dev):
        out_c = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        out_o = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
        out_w = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        match = torch.zeros(B, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if B:
            nbytes = lib.cl4_refine_scratch_bytes(B, H, W)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.cl4_pseudo_labels(_lib.ptr(gt), _lib.ptr(lab), _lib.ptr(conf), _lib.ptr(ys), _lib.ptr(xs), K,
                                             float(pseudo_thresh), _lib.ptr(_gauss_f32(sigma, dev)), sigma,
                                             MINIMUM_MASK_SIZE, _lib.ptr(out_c), _lib.ptr(out_o), _lib.ptr(out_w),
                                             _lib.ptr(match), _lib.ptr(status), B, C, H, W, _lib.ptr(scratch), nbytes,
                                             _lib.stream_ptr(dev)), "pseudo_label_generation")
    if int(status.item()) != 0:
        raise NotImplementedError(f"pseudo_label_generation_batch: more than {lib.cl4_refine_max_contours()} contours "
                                  f"of >= {MINIMUM_MASK_SIZE} px in one image")
    return out_c, out_o, out_w, match
