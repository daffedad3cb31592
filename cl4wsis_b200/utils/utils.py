"""Drop-in ``denorm`` (reference utils/utils.py:26-41) on an sm_100a kernel: it produces PAMR's image
input inside the phase-1 step (train.py:376)."""
import torch

from .. import _lib


def denorm(image, mean=(0.485, 0.456, 0.4069), std=(0.229, 0.224, 0.225)):
    """``image * std + mean`` per RGB channel on a copy; [3,H,W] or [B,3,H,W] fp32 CUDA.  The default mean keeps
    the reference's 0.4069 (utils/utils.py:26)."""
    lib = _lib.load()
    _lib.require_cuda(image, "image")
    if image.dim() == 3:
        assert image.size(0) == 3, "Expected RGB image [3xHxW]"
    elif image.dim() == 4:
        assert image.size(1) == 3, "Expected RGB image [3xHxW]"
    else:
        return image.clone()  # the reference clones and touches nothing for other ranks
    if image.dtype != torch.float32:
        raise TypeError("denorm: fp32 only on this path")
    x = image.detach().contiguous()
    out = torch.empty_like(x)
    HW = x.shape[-1] * x.shape[-2]
    planes = x.numel() // HW if HW else 0
    if planes == 0 or HW == 0:
        return out
    with torch.cuda.device(x.device):
        _lib.check(lib.cl4_denorm(_lib.ptr(x), _lib.ptr(out), planes, 3, HW, _lib.float_array(mean[:3]),
                                  _lib.float_array(std[:3]), _lib.stream_ptr(x.device)), "denorm")
    return out
