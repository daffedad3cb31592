"""Sweep time of cl4_pamr_forward_timed for several batch sizes (dev tool).  python tools/time_pamr.py [C] [H] [B...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cl4wsis_b200 as cl4
C = int(sys.argv[1]) if len(sys.argv) > 1 else 21
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Bs = [int(b) for b in sys.argv[3:]] or [1, 2, 4, 16]
T = 10
dil = [int(d) for d in os.environ.get("DIL", "1,2,4,8,12,24").split(",")]
for B in Bs:
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=T, dilations=dil)
    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.rand(B, 3, H, W, device="cuda", generator=g)
    mask = torch.rand(B, C, H, W, device="cuda", generator=g).softmax(1)
    heat = torch.rand(B, 1, H, W, device="cuda", generator=g) * 0.2
    off = torch.randn(B, 2, H, W, device="cuda", generator=g)
    n = 5
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs: a.record(); b.record()
    for _ in range(3): step.run(img, mask, heat, off)
    torch.cuda.synchronize()
    for e in evs: step.run(img, mask, heat, off, sweep_events=e)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / (n * T)
    tiles = B * ((H + 31) // 32) * ((W + 31) // 32)
    rounds = (tiles + 147) // 148
    print(f"dil={dil} B={B:3d} C={C} {H}x{W}: {ms:.4f} ms per sweep(+frame); {rounds} tile rounds -> {ms * 1e-3 * 1.965e9 / (rounds * C):.0f} cycles per (tile, class) item")
