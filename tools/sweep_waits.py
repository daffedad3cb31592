"""Where the warp-specialised sweep waits (needs a -DCL4_SWEEP_DEBUG build, CL4_SWEEP=ws)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cl4wsis_b200 as cl4
lib = cl4._lib.load()
B, C, H, W, T = 16, 21, 512, 512, 10
step = cl4.PseudoLabelStep(B, C, H, W, num_iter=T)
g = torch.Generator(device="cuda").manual_seed(1)
img = torch.rand(B, 3, H, W, device="cuda", generator=g); mask = torch.rand(B, C, H, W, device="cuda", generator=g).softmax(1)
heat = torch.rand(B, 1, H, W, device="cuda", generator=g) * 0.2; off = torch.randn(B, 2, H, W, device="cuda", generator=g)
for _ in range(4): step.run(img, mask, heat, off)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1480)()
lib.cl4_debug_sweep_waits.argtypes = [ctypes.c_void_p]
assert lib.cl4_debug_sweep_waits(buf) == 0
tot = [buf[10 * i] for i in range(148)]; prod = [buf[10 * i + 1] for i in range(148)]
cons = [[buf[10 * i + 2 + w] for w in range(8)] for i in range(148)]
mt = sum(tot) / 148
print(f"total cycles mean {mt:.0f}; producer waits for a free stage {100 * sum(prod) / sum(tot):.1f} % of the time")
print("consumer warps wait for data, % of the time, mean over CTAs:", [round(100 * sum(c[w] for c in cons) / sum(tot), 1) for w in range(8)])
print("per-CTA spread of the consumer wait share (min, max %):", round(100 * min(sum(c) / 8 / t for c, t in zip(cons, tot)), 1), round(100 * max(sum(c) / 8 / t for c, t in zip(cons, tot)), 1))
