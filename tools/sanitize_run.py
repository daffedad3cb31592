"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck) runs."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl4wsis_b200 as cl4
from cl4wsis_b200.modules import utils as mu
from cl4wsis_b200.wss import modules as wm
from cl4wsis_b200.wss.utils import peak_extract_device, smoothing

torch.manual_seed(0)
dev = "cuda"
# PAMR: fused (1 tile, 4 tiles, odd sizes), TMA path, register path
for (B, C, H, W, dil, T) in [(2, 5, 32, 32, [1, 2, 4, 8, 12], 3), (1, 3, 56, 56, [1, 2, 4, 8, 12], 2), (1, 2, 29, 37, [1, 2, 4, 8, 12, 24], 2),
                             (1, 3, 96, 80, [1, 2, 4, 8, 12, 24], 2), (1, 2, 40, 40, [1, 2, 3, 4, 5, 6, 7, 8], 2)]:
    y = cl4.PAMR(T, dil).cuda()(torch.rand(B, 3, H, W, device=dev), torch.rand(B, C, H, W, device=dev).softmax(1))
    assert torch.isfinite(y).all()
# the 4-pixel sweep (with its producer warpgroup) on the default dilation set, and the trainer's set on a large map
os.environ["CL4_SWEEP"] = "nolattice"
cl4.PAMR(2, [1, 2, 4, 8, 12, 24]).cuda()(torch.rand(1, 3, 96, 80, device=dev), torch.rand(1, 3, 96, 80, device=dev).softmax(1))
del os.environ["CL4_SWEEP"]
cl4.PAMR(2, [1, 2, 4, 8, 12]).cuda()(torch.rand(1, 3, 72, 100, device=dev), torch.rand(1, 2, 72, 100, device=dev).softmax(1))
# phase-1 producers / consumers
from cl4wsis_b200.wss import single_stage as ss
ss.phase1_pseudo_labels(torch.randn(2, 3, 64, 80, device=dev), torch.randn(2, 5, 16, 20, device=dev), torch.ones(2, 4, device=dev),
                        cl4.PAMR(3, [1, 2, 4, 8, 12]).cuda())
x = torch.rand(1, 2, 20, 33, device=dev)
for cls in (wm.LocalAffinity, wm.LocalAffinityAbs, wm.LocalAffinityCopy, wm.LocalStDev):
    cls([1, 2, 24]).cuda()(x)
heat = torch.rand(2, 3, 70, 90, device=dev)
pk = peak_extract_device(smoothing(heat), 15, 25)
peak_extract_device(heat - 0.5, 3, 7)
peak_extract_device(heat, 9, 40)
ctr = cl4.find_instance_center(heat[:1, :1], 0.5, 5)
ids = cl4.group_pixels(ctr, torch.randn(1, 2, 70, 90, device=dev))
cl4.get_instance_segmentation(torch.rand(1, 70, 90, device=dev) > 0.3, heat[:1, :1].clone(), torch.randn(1, 2, 70, 90, device=dev), 0.3, 41, None, True, 3.0)


class A:
    refine_thresh, kernel, beta, sigma = 0.3, 41, 3.0, 6


B, C, H, W = 2, 3, 70, 90
gt = torch.zeros(B, H, W, dtype=torch.long, device=dev)
gt[:, 10:40, 10:50] = 1
gt[:, 45:65, 30:80] = 2
gt[1, 5:9, 60:64] = 3
lab = torch.ones(B, C, device=dev)
off = torch.randn(B, 2, H, W, device=dev)
off[:, :, 20:25, 20:25] *= 0.1
out, st = mu.refine_label_generation_device(torch.randn(B, C + 1, H, W, device=dev), heat, off, lab, gt, 10000, A)
mu.refine_label_generation_per_contour(torch.randn(B, C + 1, H, W, device=dev), heat, off, lab, gt, 10000, A)
mu.pseudo_label_generation_batch(gt, pk, lab, 0.7, 6)
step = cl4.PseudoLabelStep(2, 4, 64, 96, num_iter=3, dilations=[1, 2, 4, 8, 12, 24], threshold=0.3, nms_kernel=41, max_centers=32)
step.run(torch.rand(2, 3, 64, 96, device=dev), torch.rand(2, 4, 64, 96, device=dev).softmax(1), torch.rand(2, 1, 64, 96, device=dev),
         torch.randn(2, 2, 64, 96, device=dev))
# round-2 paths: class-pair sweep with an odd class count and partial tiles, K > 256, the phase-2 CAM chain with the
# up-sampling inside the tile loader, get_ins_map on the device
cl4.PAMR(3, [1, 2, 4, 8, 12, 24]).cuda()(torch.rand(1, 3, 100, 132, device=dev), torch.rand(1, 5, 100, 132, device=dev).softmax(1))
peak_extract_device(heat, 5, 300)
from cl4wsis_b200.wss.utils import cam_peaks
cam_peaks(torch.randn(2, 3, 20, 24, device=dev), torch.ones(2, 3, device=dev), (70, 90))
from cl4wsis_b200.dataset.utils import get_ins_map


class V:
    val_flip, val_clean, val_thresh, val_kernel, beta, val_ignore = True, True, 0.3, 41, 3.0, True


seg = torch.randn(2, C + 1, H, W, device=dev)
seg[:, 1, 10:60, 10:50] += 6
seg[:, 2, 20:65, 55:85] += 6
get_ins_map({'seg': seg, 'center': heat[:2].contiguous(), 'offset': torch.randn(2, 2, H, W, device=dev)}, torch.ones(1, C), (H, W), dev, V)
torch.cuda.synchronize()
print("sanitize_run ok, status", int(st.item()))
