"""Second oracle / second baseline: the hot path restated with STOCK PyTorch ops, runnable on CUDA tensors.

TEST INFRASTRUCTURE.  Only tests/ and bench.py's `torch_gpu_baseline` leg import this file; the product
(cl4wsis_b200/) never does.  The reference itself cannot travel to the GPU box (/root/reference does not
exist there), so this file restates, op for op, what the reference executes when the trainer runs it on
CUDA tensors (SURVEY §8c "GPU oracle", §8d "also time the reference on the GPU via stock PyTorch"):

  pamr                    wss/modules.py:17-152   F.pad(replicate) + F.conv2d(dilation=d) with the 3x3 shift kernels,
                                                  std over the 9D samples, softmax over the 8D taps, num_iter x (m * x).sum(2)
  find_instance_center    modules/utils.py:463-502  F.threshold + F.max_pool2d + nonzero
  group_pixels            modules/utils.py:505-542  torch.norm + argmin
  peak_extract            wss/utils.py:3-25         F.max_pool2d + topk
  smoothing               wss/utils.py:28-32        F.avg_pool2d

The functions are checked against the live reference on CPU in tests/test_oracle_live_reference.py (where
/root/reference exists); on the GPU box the same code runs on CUDA tensors and pins the arithmetic of ATen's CUDA
kernels (norm, max-pool, conv) that the CPU oracle cannot see.
"""
import torch
import torch.nn.functional as F

# (dy, dx) of the eight taps in the reference's kernel order (wss/modules.py:31-40): row-major 3x3 without the centre
_TAPS = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)]


def _shift_kernels(kind, dtype, device):
    """kind 'diff': centre - neighbour (LocalAffinity, :26-40); 'copy': neighbour (LocalAffinityCopy, :67-79);
    'all9': the nine positions (LocalStDev, :86-100)."""
    if kind == "all9":
        w = torch.zeros(9, 1, 3, 3, dtype=dtype, device=device)
        for i in range(9):
            w[i, 0, i // 3, i % 3] = 1
        return w
    w = torch.zeros(8, 1, 3, 3, dtype=dtype, device=device)
    for i, (r, c) in enumerate(_TAPS):
        if kind == "diff":
            w[i, 0, 1, 1] = 1
            w[i, 0, r, c] = -1
        else:
            w[i, 0, r, c] = 1
    return w


def _local(x, kernel, dilations):
    """LocalAffinity.forward (:47-62): [B,K,H,W] -> [B,K,len(kernel)*D,H,W]."""
    B, K, H, W = x.shape
    x = x.reshape(B * K, 1, H, W)
    outs = [F.conv2d(F.pad(x, [d] * 4, mode="replicate"), kernel, dilation=d) for d in dilations]
    return torch.cat(outs, 1).view(B, K, -1, H, W)


def pamr_weights(x, dilations):
    """wss/modules.py:141-146: [B,K,H,W] -> affinity [B,1,8D,H,W]."""
    std = _local(x, _shift_kernels("all9", x.dtype, x.device), dilations).std(2, keepdim=True)
    aff = -_local(x, _shift_kernels("diff", x.dtype, x.device), dilations).abs() / (1e-8 + 0.1 * std)
    return F.softmax(aff.mean(1, keepdim=True), 2)


def pamr(x, mask, num_iter=10, dilations=(1, 2, 4, 8, 12, 24)):
    """PAMR.forward (wss/modules.py:133-152)."""
    mask = F.interpolate(mask, size=x.shape[-2:], mode="bilinear", align_corners=True)
    aff = pamr_weights(x, dilations)
    copy = _shift_kernels("copy", mask.dtype, mask.device)
    for _ in range(num_iter):
        mask = (_local(mask, copy, dilations) * aff).sum(2)
    return mask


def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=5):
    """modules/utils.py:463-494 (top_k=None): [1,1,H,W] -> [K,2] (y, x) int64."""
    h = F.threshold(ctr_hmp, threshold, -1)
    pooled = F.max_pool2d(h, kernel_size=nms_kernel, stride=1, padding=(nms_kernel - 1) // 2)
    h = torch.where(h != pooled, torch.full_like(h, -1), h)
    return torch.nonzero(h.squeeze() > 0, as_tuple=False)


def group_pixels(ctr, offsets):
    """modules/utils.py:505-542: ctr [K,2], offsets [1,2,H,W] -> [1,H,W] int64 ids in 1..K."""
    off = offsets.squeeze(0)
    H, W = off.shape[1:]
    yy = torch.arange(H, dtype=off.dtype, device=off.device).view(1, H, 1).expand(1, H, W)
    xx = torch.arange(W, dtype=off.dtype, device=off.device).view(1, 1, W).expand(1, H, W)
    loc = (torch.cat((yy, xx), 0) + off).reshape(2, H * W).transpose(1, 0)
    dist = torch.norm(ctr.unsqueeze(1) - loc.unsqueeze(0), dim=-1)
    return torch.argmin(dist, dim=0).reshape(1, H, W) + 1


def peak_extract(heat, kernel=5, K=25):
    """wss/utils.py:3-25 on device tensors: (scores [B,C,K] f32, ys i32, xs i32)."""
    B, C, H, W = heat.shape
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=(kernel - 1) // 2)
    peak = heat * (hmax == heat).float()
    s, i = torch.topk(peak.view(B, C, -1), K)
    i = i % (H * W)
    return s.float(), (i / W).int(), (i % W).int()


def smoothing(heat, kernel=3):
    """wss/utils.py:28-32."""
    return F.avg_pool2d(heat, (kernel, kernel), stride=1, padding=(kernel - 1) // 2)


def pseudo_label_step(img, mask, heat, off, num_iter, dilations, threshold, nms_kernel):
    """One step of the headline workload as the reference would run it on CUDA tensors: batched PAMR, then the
    per-image (batch-1) centre NMS + grouping loop.  Returns (refined, [ids per image])."""
    refined = pamr(img, mask, num_iter, dilations)
    ids = []
    for b in range(heat.shape[0]):
        ctr = find_instance_center(heat[b:b + 1].clone(), threshold, nms_kernel)
        ids.append(group_pixels(ctr, off[b:b + 1]) if ctr.shape[0] else torch.zeros_like(off[b:b + 1, 0]).long())
    return refined, ids
