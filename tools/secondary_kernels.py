"""The secondary kernels of the hot path on the headline shapes (B16, 512x512), each timed with CUDA events (or captured:
`ncu --set full -k regex:'center_flags|center_compact|peak_tile|peak_merge|group_pixels|weights_lattice|cam_normalize|smoothing'
python tools/secondary_kernels.py --once`).  Prints, per kernel, the algorithmic bytes and the fraction of the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cl4wsis_b200 as cl4
from cl4wsis_b200.wss import utils as wu
L = cl4._lib; lib = L.load()
once = "--once" in sys.argv
B, C, H, W = 16, 21, 512, 512
cfg = bench.WORKLOADS["voc_b16_c21_512"]
img, mask, heat, off = (t.cuda() for t in bench.synth_inputs(cfg, 0, device="cuda"))
peak, _ = bench.measured_peak_gbs()
st = L.stream_ptr()
res = {}


def timeit(name, fn, nbytes, n=20):
    if once:
        fn(); torch.cuda.synchronize(); return
    ms = bench.ev_time(fn, n=n)
    res[name] = {"us": 1e3 * ms, "algorithmic_bytes": nbytes, "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    print(f"{name:58s} {1e3 * ms:9.1f} us  {nbytes / 1e6:8.1f} MB  {res[name]['hbm_frac']:.3f} of HBM peak")


nb = lib.cl4_center_nms_scratch_bytes(B, H, W)
scr = torch.empty(nb, dtype=torch.uint8, device="cuda"); ctr = torch.empty(B, 256, 2, dtype=torch.int64, device="cuda"); cnt = torch.empty(B, dtype=torch.int32, device="cuda")
timeit("center_nms B16 512^2 k41 thr0.3 (5 gaussians per image)", lambda: L.check(lib.cl4_center_nms(L.ptr(heat), 0.3, 0.0, 41, B, H, W, L.ptr(ctr), L.ptr(cnt), 256, L.ptr(scr), nb, st), "n"), 4.0 * B * H * W)
dense = torch.rand(B, 1, H, W, device="cuda")
timeit("center_nms B16 512^2 k41 thr0.3 (uniform noise: every tile live)", lambda: L.check(lib.cl4_center_nms(L.ptr(dense), 0.3, 0.0, 41, B, H, W, L.ptr(ctr), L.ptr(cnt), 256, L.ptr(scr), nb, st), "n"), 4.0 * B * H * W)
ids = torch.empty(B, H, W, dtype=torch.int64, device="cuda")
c5 = torch.randint(0, 512, (B, 256, 2), device="cuda")
timeit("group_pixels B16 512^2 Kc=5", lambda: L.check(lib.cl4_group_pixels(L.ptr(c5), None, 5, 256, L.ptr(off), None, L.ptr(ids), B, H, W, 0, st), "g"), 16.0 * B * H * W)
timeit("group_pixels B16 512^2 Kc=200 (ALU-bound)", lambda: L.check(lib.cl4_group_pixels(L.ptr(c5), None, 200, 256, L.ptr(off), None, L.ptr(ids), B, H, W, 0, st), "g"), 16.0 * B * H * W)
# phase-2 chain: raw CAM at feature resolution -> peaks at image resolution (train.py:426-436)
g = torch.Generator(device="cuda").manual_seed(3)
yy = torch.arange(32, device="cuda").view(1, 1, 32, 1).float(); xx = torch.arange(32, device="cuda").view(1, 1, 1, 32).float()
cam = 0.3 * torch.randn(B, 20, 32, 32, device="cuda", generator=g)
for _ in range(3):
    cy = 32 * torch.rand(B, 20, 1, 1, device="cuda", generator=g); cx = 32 * torch.rand(B, 20, 1, 1, device="cuda", generator=g)
    cam += 3 * torch.rand(B, 20, 1, 1, device="cuda", generator=g) * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 6.0)
lab = (torch.rand(B, 20, device="cuda", generator=g) < 0.2).float()
small = wu.smoothing(wu.cam_normalize(cam, None, lab), 3)
up = torch.nn.functional.interpolate(small, size=(H, W), mode="bilinear", align_corners=False)
timeit("peak_extract B16 C20 512^2 k15 K25 (materialised map)", lambda: wu.peak_extract_device(up, 15, 25), 4.0 * B * 20 * H * W)
timeit("peak_extract_upsampled 32^2 -> 512^2 (fused loader)", lambda: wu.peak_extract_device(small, 15, 25, upsample_to=(H, W)), 4.0 * B * 20 * 32 * 32)
timeit("cam_peaks: cam_normalize+smoothing+upsample+peaks (4 launches)", lambda: wu.cam_peaks(cam, lab, (H, W)), 4.0 * B * 20 * 32 * 32)
timeit("F.interpolate + peak_extract (what the fused loader replaces)", lambda: wu.peak_extract_device(torch.nn.functional.interpolate(small, size=(H, W), mode="bilinear", align_corners=False), 15, 25), 3 * 4.0 * B * 20 * H * W)
mod = cl4.PAMR(1, [1, 2, 4, 8, 12, 24]).cuda()
timeit("PAMR num_iter=1 (pad + weights_lattice + one sweep)", lambda: mod(img, mask), 4.0 * B * H * W * (3 + 48 + 48 + 2 * C))
if not once:
    print(json.dumps(res))
