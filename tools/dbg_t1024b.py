import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import cl4wsis_b200 as cl4
import oracle
rng = np.random.default_rng(3 * 7 + 1024)
B,C,H,W = 1,3,1024,1024
dil=[1,2,4,8,12,24]
x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
T = int(os.environ.get("T", "1"))
want = oracle.pamr(x, m, T, dil)
for rep in range(3):
    got = cl4.PAMR(T, dil).cuda()(torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda()).cpu().numpy()
    bad = np.argwhere(~np.isclose(got, want, rtol=1e-4, atol=1e-6))
    print("T", T, "bad", len(bad))
    if len(bad):
        ys, xs = bad[:,2], bad[:,3]
        tiles = np.unique((ys//32)*32 + xs//32)
        print(" classes", np.bincount(bad[:,1], minlength=3), "n tiles", len(tiles), "tiles%148", np.unique(tiles % 148)[:30])
        print(" y%32 hist", np.bincount(ys%32, minlength=32))
        print(" x%32 hist", np.bincount(xs%32, minlength=32))
