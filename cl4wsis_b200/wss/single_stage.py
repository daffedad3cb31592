"""Drop-in ``pseudo_gtmask`` (reference wss/single_stage.py:18-40) and the fused phase-1 pseudo-label step
around PAMR (train.py:372-385) on sm_100a kernels."""
import torch

from .. import _lib


def _f32c(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: fp32 only on this path")
    return t.detach().contiguous()


def pseudo_gtmask(mask, ambiguous=True, cutoff_top=0.6, cutoff_bkg=0.6, cutoff_low=0.2, eps=1e-8, old_classes=16):
    """Convert continuous mask into binary mask: [B,C,h,w] fp32 CUDA -> 0/1 floats of the same shape.
    ``eps`` and ``old_classes`` are unused, as in the reference."""
    lib = _lib.load()
    m = _f32c(mask, "mask")
    bs, c, h, w = m.shape
    out = torch.empty_like(m)
    if m.numel() == 0:
        return out
    with torch.cuda.device(m.device):
        thr = torch.empty((bs, c), dtype=torch.float32, device=m.device)
        _lib.check(lib.cl4_pseudo_gtmask(_lib.ptr(m), None, None, _lib.ptr(out), _lib.ptr(thr), bs, c, h * w,
                                         float(cutoff_top), float(cutoff_bkg), float(cutoff_low), 1 if ambiguous else 0,
                                         _lib.stream_ptr(m.device)), "pseudo_gtmask")
    return out


def softmax_channels(x):
    """``x.softmax(dim=1)`` for [B,C,h,w] (train.py:372-373)."""
    lib = _lib.load()
    x = _f32c(x, "x")
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    B, C = x.shape[:2]
    with torch.cuda.device(x.device):
        _lib.check(lib.cl4_softmax_channels(_lib.ptr(x), _lib.ptr(out), B, C, x.numel() // (B * C),
                                            _lib.stream_ptr(x.device)), "softmax_channels")
    return out


def denorm_resize(images, size, mean=(0.485, 0.456, 0.4069), std=(0.229, 0.224, 0.225)):
    """``F.interpolate(denorm(images), size, mode="bilinear", align_corners=True)`` in one pass (train.py:376-378)."""
    lib = _lib.load()
    x = _f32c(images, "images")
    assert x.dim() == 4 and x.size(1) == 3, "Expected RGB image [3xHxW]"
    B, K, Hi, Wi = x.shape
    h, w = int(size[0]), int(size[1])
    out = torch.empty((B, K, h, w), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        _lib.check(lib.cl4_denorm_resize_ac(_lib.ptr(x), _lib.ptr(out), B, K, Hi, Wi, h, w, _lib.float_array(mean[:3]),
                                            _lib.float_array(std[:3]), _lib.stream_ptr(x.device)), "denorm_resize")
    return out


PHASE1_LAUNCHES = 2  # kernels per fused phase1_pseudo_labels call (prologue + on-chip PAMR with epilogue); 6 otherwise


def _fused_applicable(affinity, h, w):
    dil = [int(d) for d in affinity.aff_x.dilations]
    return (affinity.num_iter >= 1 and 1 <= len(dil) <= 6 and max(dil) <= 24 and min(dil) >= 1 and h <= 64 and w <= 64)


def phase1_pseudo_labels(images, int_masks, l1h, affinity=None, cutoff_top=0.6, cutoff_bkg=0.7, cutoff_low=0.2,
                         mean=(0.485, 0.456, 0.4069), std=(0.229, 0.224, 0.225)):
    """train.py:372-385: softmax over classes, denorm + bilinear shrink of the images to the masks' resolution, PAMR
    (``affinity``: a ``cl4wsis_b200.PAMR``; None = ``use_aff`` off), gating of the foreground planes by the image-level
    labels ``l1h`` [B,C-1] and ``pseudo_gtmask(..., ambiguous=True, 0.6, 0.7, 0.2)``.
    Returns (int_masks_soft gated [B,C,h,w], pseudo_gt_seg [B,C,h,w]).  Both are DETACHED results (the reference detaches
    the PAMR input and the pseudo labels, train.py:379,385): they carry no gradient and must not stand in for the
    autograd-tracked ``int_masks_orig`` of train.py:372 that the losses at :386-411 differentiate.

    Feature-resolution maps (<= 64 x 64, up to six dilations <= 24: the trainer's regime) take ``cl4_phase1_pseudo_labels``:
    TWO launches for the whole step.  Larger maps run the same pieces as six launches."""
    lib = _lib.load()
    if affinity is not None and int_masks.dim() == 4 and _fused_applicable(affinity, *int_masks.shape[-2:]):
        x = _f32c(images, "images")
        m = _f32c(int_masks, "int_masks")
        assert x.dim() == 4 and x.size(1) == 3, "Expected RGB image [3xHxW]"
        B, C, h, w = m.shape
        lab = _f32c(l1h.to(torch.float32), "l1h")
        assert lab.shape == (B, C - 1), "l1h must be [B, C-1] (train.py:382)"
        dil = [int(d) for d in affinity.aff_x.dilations]
        soft = torch.empty_like(m)
        pseudo = torch.empty_like(m)
        if m.numel() == 0:
            return soft, pseudo
        with torch.cuda.device(m.device):
            nbytes = lib.cl4_phase1_scratch_bytes(B, C, h, w, len(dil))
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=m.device)
            _lib.check(lib.cl4_phase1_pseudo_labels(_lib.ptr(x), _lib.ptr(m), _lib.ptr(lab), _lib.float_array(mean[:3]),
                                                    _lib.float_array(std[:3]), _lib.int_array(dil), len(dil),
                                                    int(affinity.num_iter), float(cutoff_top), float(cutoff_bkg),
                                                    float(cutoff_low), _lib.ptr(soft), _lib.ptr(pseudo), _lib.ptr(scratch),
                                                    nbytes, B, C, x.shape[2], x.shape[3], h, w, _lib.stream_ptr(m.device)),
                       "phase1_pseudo_labels")
        return soft, pseudo
    return phase1_pseudo_labels_unfused(images, int_masks, l1h, affinity, cutoff_top, cutoff_bkg, cutoff_low)


def phase1_pseudo_labels_unfused(images, int_masks, l1h, affinity=None, cutoff_top=0.6, cutoff_bkg=0.7, cutoff_low=0.2):
    """The same step as separate launches (softmax, denorm + shrink, PAMR's own kernels, gating + thresholds, pseudo labels):
    any map size and dilation set."""
    lib = _lib.load()
    soft = softmax_channels(int_masks)
    if affinity is not None:
        im = denorm_resize(images, int_masks.shape[-2:])
        soft = affinity(im, soft)
    B, C, h, w = soft.shape
    lab = _f32c(l1h.to(torch.float32), "l1h")
    assert lab.shape == (B, C - 1), "l1h must be [B, C-1] (train.py:382)"
    pseudo = torch.empty_like(soft)
    if soft.numel() == 0:
        return soft, pseudo
    with torch.cuda.device(soft.device):
        thr = torch.empty((B, C), dtype=torch.float32, device=soft.device)
        # gated in place: cl4_pseudo_gtmask allows gated_out == mask (each element is read, then written, by one thread)
        _lib.check(lib.cl4_pseudo_gtmask(_lib.ptr(soft), _lib.ptr(lab), _lib.ptr(soft), _lib.ptr(pseudo), _lib.ptr(thr),
                                         B, C, h * w, float(cutoff_top), float(cutoff_bkg), float(cutoff_low), 1,
                                         _lib.stream_ptr(soft.device)), "phase1_pseudo_labels")
    return soft, pseudo
