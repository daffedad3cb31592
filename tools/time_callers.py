"""Times the widened rows on the GPU next to the CPU oracle (dev tool): refine_label_generation,
smoothing -> peak_extract -> pseudo_label_generation, at B16 / 20 classes / 512x512."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl4wsis_b200 as cl4  # noqa: E402
from cl4wsis_b200.modules import utils as mu  # noqa: E402
from cl4wsis_b200.wss.utils import peak_extract_device, smoothing  # noqa: E402


class Args:
    refine_thresh, kernel, beta, sigma = 0.3, 41, 3.0, 6


def scene(B, C, H, W, seed=0, blobs=8):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    seg = rng.standard_normal((B, C + 1, H, W)).astype(np.float32)
    heat = (0.05 * rng.random((B, C, H, W))).astype(np.float32)
    off = (0.3 * rng.standard_normal((B, 2, H, W)) + 40).astype(np.float32)
    gt = np.zeros((B, H, W), np.int64)
    lab = np.zeros((B, C), np.float32)
    for b in range(B):
        for _ in range(blobs):
            cls = int(rng.integers(0, C)); cy, cx = int(rng.integers(20, H - 20)), int(rng.integers(20, W - 20))
            ry, rx = int(rng.integers(10, 80)), int(rng.integers(10, 80))
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
            gt[b][m] = cls + 1
            lab[b, cls] = 1
            heat[b, cls] = np.maximum(heat[b, cls], rng.uniform(0.4, 0.95) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72).astype(np.float32))
            off[b, 0][m] = (cy - yy)[m] + 0.3 * rng.standard_normal(int(m.sum()))
            off[b, 1][m] = (cx - xx)[m] + 0.3 * rng.standard_normal(int(m.sum()))
        for c in range(C + 1):
            seg[b, c][gt[b] == c] += 3
    return seg, heat, off, lab, gt


def gpu_time(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, C, H, W = 16, 20, 512, 512
    seg, heat, off, lab, gt = scene(B, C, H, W)
    d = [torch.from_numpy(a).cuda() for a in (seg, heat, off, lab, gt)]
    ms = gpu_time(lambda: mu.refine_label_generation_device(*d, 10000, Args)[0])
    ms_sync = gpu_time(lambda: mu.refine_label_generation(*d, 10000, Args))
    t0 = time.perf_counter()
    mu.refine_label_generation_per_contour(*[t[:2] for t in d], 10000, Args)
    torch.cuda.synchronize()
    pc = (time.perf_counter() - t0) * 1e3 / 2
    print(f"refine_label_generation B{B} C{C} {H}x{W}: batched {ms:.3f} ms ({B / ms * 1e3:.0f} img/s), with status sync {ms_sync:.3f} ms; "
          f"per-contour path {pc:.1f} ms/img")
    if "--cpu" in sys.argv:
        import oracle
        t0 = time.perf_counter()
        oracle.labelgen.refine_label_generation(seg[:2], heat[:2], off[:2], lab[:2], gt[:2], 10000)
        print(f"  CPU oracle (numpy + OpenCV + C): {(time.perf_counter() - t0) / 2 * 1e3:.1f} ms/img")
    cam = torch.from_numpy(heat).cuda()

    def phase2():
        sm = smoothing(cam)
        pk = peak_extract_device(sm, 15, 25)
        return mu.pseudo_label_generation_batch(d[4], pk, d[3], 0.7, 6)
    ms = gpu_time(phase2)
    print(f"smoothing + peak_extract + pseudo_label_generation B{B} C{C} {H}x{W}: {ms:.3f} ms ({B / ms * 1e3:.0f} img/s)")


if __name__ == "__main__":
    main()
