"""Drop-in ``peak_extract`` and ``smoothing`` (reference wss/utils.py:3-32) on sm_100a kernels."""
import torch

from .. import _lib


def peak_extract_device(heat, kernel=5, K=25, upsample_to=None):
    """Device-resident variant: returns (scores f32, ys i32, xs i32) CUDA tensors [B,C,K].
    ``upsample_to=(H, W)``: the peaks of ``F.interpolate(heat, (H, W), mode="bilinear", align_corners=False)``
    (train.py:431-436) without materialising the up-sampled map (``cl4_peak_extract_upsampled``)."""
    lib = _lib.load()
    _lib.require_cuda(heat, "heat")
    if heat.dim() != 4:
        raise ValueError("peak_extract expects heat [B,C,H,W]")
    heat = heat.detach()
    if heat.dtype != torch.float32:
        heat = heat.float()
    heat = heat.contiguous()
    B, C, H, W = heat.shape
    h = w = 0
    if upsample_to is not None:
        h, w = H, W
        H, W = int(upsample_to[0]), int(upsample_to[1])
    if kernel % 2 == 0:
        # the reference fails at `hmax == heat` (wss/utils.py:11): an even kernel shrinks the pooled map
        raise RuntimeError(f"peak_extract: even kernel {kernel} makes max_pool2d's output smaller than heat")
    if K > H * W:
        raise RuntimeError("selected index k out of range")  # torch.topk's message
    dev = heat.device
    with torch.cuda.device(dev):
        scores = torch.empty((B, C, K), dtype=torch.float32, device=dev)
        ys = torch.empty((B, C, K), dtype=torch.int32, device=dev)
        xs = torch.empty((B, C, K), dtype=torch.int32, device=dev)
        if B * C == 0:
            return scores, ys, xs
        nbytes = lib.cl4_peak_extract_scratch_bytes(B, C, H, W, int(kernel), int(K))
        scratch = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        if upsample_to is None:
            _lib.check(lib.cl4_peak_extract(_lib.ptr(heat), _lib.ptr(scores), _lib.ptr(ys), _lib.ptr(xs),
                                            _lib.ptr(scratch), nbytes, B, C, H, W, int(kernel), int(K),
                                            _lib.stream_ptr(dev)), "peak_extract")
        else:
            _lib.check(lib.cl4_peak_extract_upsampled(_lib.ptr(heat), h, w, _lib.ptr(scores), _lib.ptr(ys), _lib.ptr(xs),
                                                      _lib.ptr(scratch), nbytes, B, C, H, W, int(kernel), int(K),
                                                      _lib.stream_ptr(dev)), "peak_extract_upsampled")
    return scores, ys, xs


def cam_normalize(cam, size, label):
    """``PeakGenerator.cam_normalize`` (wss/modules.py:425-434) as a function: relu, gating by the image-level labels
    [B,C], bilinear resize to ``size`` (align_corners=False; ``size == cam.shape[-2:]``, the trainer's call, is the
    identity), division by (plane maximum + 1e-5).  cam [B,C,h,w] fp32 CUDA -> [B,C,*size]."""
    lib = _lib.load()
    _lib.require_cuda(cam, "cam")
    x = cam.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    B, C, h, w = x.shape
    hs, ws = (h, w) if size is None else (int(size[0]), int(size[1]))
    lab = label.detach().to(device=x.device, dtype=torch.float32).contiguous()
    if lab.shape != (B, C):
        raise RuntimeError(f"cam_normalize: label must be [B, C] = [{B}, {C}], got {tuple(lab.shape)}")
    out = torch.empty((B, C, hs, ws), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.cl4_cam_normalize(_lib.ptr(x), _lib.ptr(lab), _lib.ptr(out), B, C, h, w, hs, ws,
                                         _lib.stream_ptr(x.device)), "cam_normalize")
    return out


def cam_peaks(cam, label, image_size, smooth_kernel=3, kernel=15, K=25):
    """The phase-2 chain of train.py:426-436 on the raw CAM of the peak generator (``x`` at wss/modules.py:409):
    cam_normalize -> smoothing -> bilinear up-sampling to the image size -> peak_extract, in four launches and without the
    [B,C,H,W] up-sampled map ever touching HBM.  Returns device tensors (scores f32, ys i32, xs i32) [B,C,K]."""
    return peak_extract_device(smoothing(cam_normalize(cam, None, label), smooth_kernel), kernel, K, upsample_to=image_size)


def peak_extract(heat, kernel=5, K=25):
    """Reference signature and return types: three numpy arrays [B,C,K] (f32, i32, i32),
    sorted by score; ties broken by the lower flat index (the reference leaves tie order
    to torch.topk)."""
    scores, ys, xs = peak_extract_device(heat, kernel, K)
    return scores.cpu().numpy(), ys.cpu().numpy(), xs.cpu().numpy()


def smoothing(heat, kernel=3):
    """wss/utils.py:28-32: k x k average pooling, stride 1, zero padding (padded cells counted).
    heat [B,C,H,W] (or any [...,H,W]) fp32 CUDA -> same shape."""
    lib = _lib.load()
    _lib.require_cuda(heat, "heat")
    if kernel % 2 == 0:
        raise NotImplementedError(f"smoothing: even kernel {kernel} shrinks the map in the reference; not supported")
    x = heat.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    H, W = x.shape[-2:]
    out = torch.empty_like(x)
    planes = x.numel() // (H * W) if x.numel() else 0
    with torch.cuda.device(x.device):
        _lib.check(lib.cl4_smoothing(_lib.ptr(x), _lib.ptr(out), planes, H, W, int(kernel), _lib.stream_ptr(x.device)),
                   "smoothing")
    return out
