import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cl4wsis_b200 as cl4
B, H, W = 16, 512, 512
yy = torch.arange(H, device="cuda").view(1, 1, H, 1).float(); xx = torch.arange(W, device="cuda").view(1, 1, 1, W).float()
cam2 = torch.zeros(B, 20, H, W, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
for _ in range(4):
    cy = torch.randint(0, H, (B, 20, 1, 1), device="cuda", generator=g).float(); cx = torch.randint(0, W, (B, 20, 1, 1), device="cuda", generator=g).float()
    cam2 = torch.maximum(cam2, torch.rand(B, 20, 1, 1, device="cuda", generator=g) * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 800.0))
for _ in range(3): cl4.wss.utils.peak_extract_device(cam2, 15, 25)
torch.cuda.synchronize()
