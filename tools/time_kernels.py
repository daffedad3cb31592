"""Times individual C-ABI entry points with CUDA events (dev tool; not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cl4wsis_b200 as cl4
L = cl4._lib; lib = L.load()
B, C, H, W = 16, 21, 512, 512
dil = L.int_array([1, 2, 4, 8, 12, 24])
img = torch.rand(B, 3, H, W, device="cuda")
w = torch.empty(B * 48 * H * W + 1024 * 1024, device="cuda")
heat = torch.rand(B, 1, H, W, device="cuda")
def timeit(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n:8.3f} ms")
st = L.stream_ptr()
timeit("pamr_weights B16 512^2 D6 (planar)", lambda: L.check(lib.cl4_pamr_weights(L.ptr(img), L.ptr(w), B, 3, H, W, dil, 6, st), "w"))
nb = lib.cl4_center_nms_scratch_bytes(B, H, W)
scr = torch.empty(nb, dtype=torch.uint8, device="cuda"); ctr = torch.empty(B, 256, 2, dtype=torch.int64, device="cuda"); cnt = torch.empty(B, dtype=torch.int32, device="cuda")
timeit("center_nms B16 512^2 k41", lambda: L.check(lib.cl4_center_nms(L.ptr(heat), 0.3, 0.0, 41, B, H, W, L.ptr(ctr), L.ptr(cnt), 256, L.ptr(scr), nb, st), "n"))
cam = torch.rand(B, 20, H, W, device="cuda")
timeit("peak_extract B16 C20 512^2 k15 K25 (uniform noise)", lambda: cl4.wss.utils.peak_extract_device(cam, 15, 25))
yy = torch.arange(H, device="cuda").view(1, 1, H, 1).float(); xx = torch.arange(W, device="cuda").view(1, 1, 1, W).float()
cam2 = torch.zeros(B, 20, H, W, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
for _ in range(4):
    cy = torch.randint(0, H, (B, 20, 1, 1), device="cuda", generator=g).float(); cx = torch.randint(0, W, (B, 20, 1, 1), device="cuda", generator=g).float()
    cam2 = torch.maximum(cam2, torch.rand(B, 20, 1, 1, device="cuda", generator=g) * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 800.0))
cam2 = torch.nn.functional.avg_pool2d(cam2, 3, 1, 1)
timeit("peak_extract B16 C20 512^2 k15 K25 (CAM-like blobs)", lambda: cl4.wss.utils.peak_extract_device(cam2, 15, 25))
off = torch.randn(B, 2, H, W, device="cuda"); ids = torch.empty(B, H, W, dtype=torch.int64, device="cuda")
c5 = torch.randint(0, 512, (B, 256, 2), device="cuda")
timeit("group_pixels B16 512^2 Kc=5", lambda: L.check(lib.cl4_group_pixels(L.ptr(c5), None, 5, 256, L.ptr(off), None, L.ptr(ids), B, H, W, 0, st), "g"))
timeit("group_pixels B16 512^2 Kc=200", lambda: L.check(lib.cl4_group_pixels(L.ptr(c5), None, 200, 256, L.ptr(off), None, L.ptr(ids), B, H, W, 0, st), "g"))
