"""phase1_pseudo_labels at the trainer's feature resolutions (for ncu launch lists): B16 C21 32x32 and B16 C81 56x56."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl4wsis_b200 as cl4
from cl4wsis_b200.wss import single_stage as ss
mod = cl4.PAMR(10, [1, 2, 4, 8, 12]).cuda()
for (C, h, Hi) in [(21, 32, 512), (81, 56, 448)]:
    g = torch.Generator(device="cuda").manual_seed(7)
    images = torch.randn((16, 3, Hi, Hi), generator=g, device="cuda")
    logits = 3 * torch.randn((16, C, h, h), generator=g, device="cuda")
    l1h = (torch.rand((16, C - 1), generator=g, device="cuda") < 0.1).float()
    for _ in range(4):
        soft, pseudo = ss.phase1_pseudo_labels(images, logits, l1h, mod)
    torch.cuda.synchronize()
    print("ok", C, h, float(soft.sum()), float(pseudo.sum()))
