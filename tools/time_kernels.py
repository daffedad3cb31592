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
timeit("peak_extract B16 C20 512^2 k15 K25", lambda: cl4.wss.utils.peak_extract_device(cam, 15, 25))
