import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import cl4wsis_b200 as cl4
import oracle
rng = np.random.default_rng(3 * 7 + 1024)
B,C,H,W = 1,3,1024,1024
dil=[1,2,4,8,12,24]
x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
want = oracle.pamr(x, m, 10, dil)
for rep in range(10):
    got = cl4.PAMR(10, dil).cuda()(torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda()).cpu().numpy()
    bad = np.argwhere(~np.isclose(got, want, rtol=1e-4, atol=1e-6))
    print("lib", os.environ.get("CL4_LIB"), "bad", len(bad))
    if len(bad):
        ys, xs = bad[:,2], bad[:,3]
        print(" classes", np.unique(bad[:,1]), "tile rows", np.unique(ys//32), "tile cols", np.unique(xs//32), " y%32", np.unique(ys%32)[:40], "x%32", np.unique(xs%32)[:40])
import ctypes
lib = cl4._lib.load()
if hasattr(lib, "cl4_debug_duo"):
    buf = (ctypes.c_ulonglong * 16)()
    lib.cl4_debug_duo(buf)
    print("dbg counters", list(buf))
