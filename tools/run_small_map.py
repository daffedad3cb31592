"""One PAMR call at the trainer's feature resolution (for ncu captures of the fused kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl4wsis_b200 as cl4
x = torch.rand(16, 3, 32, 32, device="cuda"); m = torch.rand(16, 21, 32, 32, device="cuda").softmax(1)
mod = cl4.PAMR(10, [1, 2, 4, 8, 12]).cuda()
for _ in range(4):
    y = mod(x, m)
torch.cuda.synchronize()
print("ok", float(y.sum()))
