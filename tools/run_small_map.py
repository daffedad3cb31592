"""PAMR calls at the trainer's feature resolutions (for ncu captures of the on-chip kernel): B16 C21 32x32 and B16 C81 56x56,
dilations [1,2,4,8,12], 10 iterations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl4wsis_b200 as cl4
mod = cl4.PAMR(10, [1, 2, 4, 8, 12]).cuda()
for (C, h) in [(21, 32), (81, 56)]:
    x = torch.rand(16, 3, h, h, device="cuda"); m = torch.rand(16, C, h, h, device="cuda").softmax(1)
    for _ in range(4):
        y = mod(x, m)
    torch.cuda.synchronize()
    print("ok", C, h, float(y.sum()))
