#!/usr/bin/env python
"""Headline benchmark: pseudo-labelled images/sec for PAMR + centre-NMS + grouping at 512x512.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One step = one pass of the hot path over one batch of synthetic input per GPU
(BASELINE.json configs[1]: batch 16, 21 classes, 512x512, dilations [1,2,4,8,12,24], 10
iterations, nms kernel 41, threshold 0.3).  Prints ONE JSON line (rank 0).

  value        whole-job images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-buffer pipeline: pinned host -> device copies of the
               inputs and device -> host copies of the refined masks + ids inside the timed region
  roofline     the PAMR propagation sweep (dominant kernel): algorithmic bytes per launch
               / mean launch time measured with CUDA events inside the timed region
  cpu_baseline the CPU oracle (a C/OpenMP port of the reference path) on this box's host cores,
               on a bounded sample of the same workload (rank 0, N=1 only)

  sustained    when the K timed steps take < 2 s: a second, >= 2 s block of the same step with NVML clocks
               sampled every 20 ms (power / thermal steady state); `value` stays the K-step figure
  small_map    the trainer's real regime (train.py:376-379): phase1_pseudo_labels at 32x32 / 56x56
  extra_workloads  short runs of BASELINE configs 3 and 4 (N=1 only)
  sbd_pass     BASELINE config 5: one 10 582-image pass sharded over the N ranks (device-resident, rolled synthetic batches)
  torch_gpu_baseline  the same step restated with stock PyTorch ops on this GPU (oracle/torch_ref.py)

`--impl reference` times the oracle port alone (the reference itself is Python and does not
travel to the GPU box).
"""
import argparse
import hashlib
import json
import math
import os
import statistics
import sys
import threading
import time

# stdout carries exactly one JSON line.  NCCL writes its version banner (and NCCL_DEBUG output) straight to file descriptor 1,
# so descriptor 1 points at stderr while the benchmark runs and is restored for the final line (_emit).
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_STDOUT_FD = None


def _divert_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: VOC 15-5 phase-2 pseudo-label step on one B200
    "voc_b16_c21_512": dict(B=16, C=21, H=512, W=512, dil=[1, 2, 4, 8, 12, 24], T=10, Kc=5, nms=41, thr=0.3),
    # configs[2]: COCO-to-VOC, 81 classes
    "coco_b16_c81_512": dict(B=16, C=81, H=512, W=512, dil=[1, 2, 4, 8, 12, 24], T=10, Kc=5, nms=41, thr=0.3),
    # configs[3]: high-res, dense instances
    "hires_b16_c21_1024": dict(B=16, C=21, H=1024, W=1024, dil=[1, 2, 4, 8, 12, 24], T=10, Kc=200, nms=41, thr=0.3),
    # tiny, for smoke runs
    "tiny": dict(B=2, C=5, H=64, W=64, dil=[1, 2, 4, 8, 12, 24], T=10, Kc=3, nms=41, thr=0.3),
}
METRIC = "pseudo-labelled images/sec (PAMR+grouping, 512^2)"
UNIT = "images/s"


# ----------------------------------------------------------------------------- synthetic inputs
def synth_inputs(cfg, first_image=0, n_images=None, device="cpu"):
    """SURVEY §8d synthetic inputs.  Every image has its own generator seeded 1234 + its GLOBAL index, so a batch does
    not depend on how the images are sharded over ranks (rank r of a weak-scaling run owns images r*B .. r*B+B-1):
    natural-like 8-bit image, dense softmax masks (what the trainer feeds PAMR: train.py:373-379), gaussian centre
    heat-map with Kc planted centres (sigma 6), offsets to the nearest centre + N(0,1)."""
    B = cfg["B"] if n_images is None else n_images
    C, H, W, Kc = cfg["C"], cfg["H"], cfg["W"], cfg["Kc"]
    dev = torch.device(device)
    F = torch.nn.functional
    yy = torch.arange(H, dtype=torch.float32, device=dev).view(1, H, 1)
    xx = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, W)
    img = torch.empty((B, 3, H, W), device=dev)
    mask = torch.empty((B, C, H, W), device=dev)
    heat = torch.zeros((B, 1, H, W), device=dev)
    off = torch.empty((B, 2, H, W), device=dev)
    for b in range(B):
        g = torch.Generator(device=dev).manual_seed(1234 + first_image + b)
        lo = torch.randint(0, 256, (1, 3, H // 8, W // 8), generator=g, device=dev).float()
        img[b] = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)[0].round().clamp(0, 255) / 255.0
        mlo = torch.randn((1, C, H // 8, W // 8), generator=g, device=dev)
        mask[b] = F.interpolate(3.0 * mlo, size=(H, W), mode="bilinear", align_corners=False)[0].softmax(0)
        cy = torch.randint(0, H, (Kc,), generator=g, device=dev).float().view(Kc, 1, 1)
        cx = torch.randint(0, W, (Kc,), generator=g, device=dev).float().view(Kc, 1, 1)
        amp = (0.5 + 0.5 * torch.rand((Kc,), generator=g, device=dev)).view(Kc, 1, 1)
        best = torch.full((H, W), float("inf"), device=dev)
        ny = torch.zeros((H, W), device=dev)
        nx = torch.zeros((H, W), device=dev)
        for k0 in range(0, Kc, 32):                                   # chunks of centres: bounded memory at Kc = 200, 1024^2
            d2 = (yy - cy[k0:k0 + 32]) ** 2 + (xx - cx[k0:k0 + 32]) ** 2  # [k,H,W]
            heat[b, 0] = torch.maximum(heat[b, 0], (amp[k0:k0 + 32] * torch.exp(-d2 / (2 * 6.0 * 6.0))).amax(0))
            dmin, near = d2.min(0)
            upd = dmin < best
            best = torch.where(upd, dmin, best)
            ny = torch.where(upd, cy[k0:k0 + 32].view(-1)[near], ny)
            nx = torch.where(upd, cx[k0:k0 + 32].view(-1)[near], nx)
            del d2
        off[b, 0] = ny - yy[0].expand(H, W) + torch.randn((H, W), generator=g, device=dev)
        off[b, 1] = nx - xx[0].expand(H, W) + torch.randn((H, W), generator=g, device=dev)
    return img.contiguous(), mask.contiguous(), heat.contiguous(), off.contiguous()


def bench_config(workload, cfg):
    """The `config` object of the JSON line -- the same dict, key for key, in both arms (ours and --impl reference)."""
    return {"workload": workload, "B_per_gpu": cfg["B"], "C": cfg["C"], "H": cfg["H"], "W": cfg["W"], "dilations": cfg["dil"],
            "num_iter": cfg["T"], "nms_kernel": cfg["nms"], "threshold": cfg["thr"], "centres_per_image": cfg["Kc"],
            "mask": "dense softmax over all classes", "l2": "inputs + scratch per step exceed L2 (no flush needed)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU oracle timing
def time_oracle(cfg, n_images, first_image, repeats=1, threads=None):
    """The oracle port on host cores over `n_images` images of the workload -> images/s."""
    import oracle as orc
    orc.build()
    if threads:
        orc.set_num_threads(threads)
    img, mask, heat, off = (t.numpy() for t in synth_inputs(cfg, first_image, n_images))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.pamr(img, mask, cfg["T"], cfg["dil"])
        for b in range(n_images):
            ctr = orc.find_instance_center(heat[b:b + 1], cfg["thr"], cfg["nms"])
            if ctr.shape[0]:
                orc.group_pixels(ctr, off[b:b + 1])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_images / best, best, orc.num_threads()


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args, cfg, rank, world):
    """The CPU arm: the oracle port on all host threads.  A step is the workload's batch (B images) whenever K + W such
    steps end within about three minutes; otherwise each step is a bounded sample of n < B images of the same batch
    (said in cpu_baseline.sample).  `config` is the same dict as in our arm."""
    if rank != 0:
        return
    thr = host_threads()
    B = cfg["B"]
    ips, _, _ = time_oracle(cfg, 1, 0, threads=thr)            # builds the oracle, pages it in, measures the rate
    budget_s = 150.0
    n_img = args.ref_images if args.ref_images > 0 else max(1, min(B, int(budget_s * ips / max(args.steps + args.warmup, 1))))
    for i in range(args.warmup):
        time_oracle(cfg, n_img, 0, threads=thr)
    t_total, cores = 0.0, None
    for i in range(args.steps):
        _, dt, cores = time_oracle(cfg, n_img, (i * B) % 4096, threads=thr)
        t_total += dt
    value = args.steps * n_img / t_total
    sample = (f"{n_img} of the {B} images of a {args.workload} batch per step x {args.steps} steps, oracle/cl4_oracle.c "
              f"(C + OpenMP), {t_total:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps * (B / n_img), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "images_per_step": n_img,
        "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; this is oracle/cl4_oracle.c "
                "(C + OpenMP port, ~12x faster than the reference's own torch-CPU path measured in SURVEY §6); "
                "ms_per_step is scaled to the full batch of B images",
    }
    _emit(line)


# ----------------------------------------------------------------------------- widened rows
def time_callers(cfg, dev, seed):
    """refine_label_generation and smoothing -> peak_extract -> pseudo_label_generation for one batch of
    the workload's shape (20 fg classes), CUDA-event timed; reported next to the headline, not part of it."""
    from cl4wsis_b200.modules import utils as mu
    from cl4wsis_b200.wss.utils import peak_extract_device, smoothing
    B, H, W, C = cfg["B"], cfg["H"], cfg["W"], 20
    g = torch.Generator(device=dev).manual_seed(seed)
    yy = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1)
    xx = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
    gt = torch.zeros((B, H, W), dtype=torch.int64, device=dev)
    heat = 0.05 * torch.rand((B, C, H, W), generator=g, device=dev)
    off = 0.3 * torch.randn((B, 2, H, W), generator=g, device=dev) + 40
    lab = torch.zeros((B, C), device=dev)
    for _ in range(8):  # 8 elliptic instances per image
        cls = torch.randint(0, C, (B,), generator=g, device=dev)
        cy = torch.randint(20, H - 20, (B, 1, 1), generator=g, device=dev).float()
        cx = torch.randint(20, W - 20, (B, 1, 1), generator=g, device=dev).float()
        ry = torch.randint(10, max(11, H // 6), (B, 1, 1), generator=g, device=dev).float()
        rx = torch.randint(10, max(11, W // 6), (B, 1, 1), generator=g, device=dev).float()
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        gt = torch.where(m, (cls + 1).view(B, 1, 1), gt)
        bump = 0.9 * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72)
        for b in range(B):
            lab[b, cls[b]] = 1
            heat[b, cls[b]] = torch.maximum(heat[b, cls[b]], bump[b])
        off[:, 0] = torch.where(m, cy - yy, off[:, 0])
        off[:, 1] = torch.where(m, cx - xx, off[:, 1])
    seg = torch.randn((B, C + 1, H, W), generator=g, device=dev)
    seg.scatter_add_(1, gt[:, None], torch.full((B, 1, H, W), 3.0, device=dev))

    class A:
        refine_thresh, kernel, beta, sigma = 0.3, 41, 3.0, 6

    def ev_time(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out, status = mu.refine_label_generation_device(seg, heat, off, lab, gt, 10000, A)
    res = {"shape": f"B{B} C{C} {H}x{W}", "refine_status": int(status.item()),
           "refine_weighted_px": int((out["weight"] > 0).sum())}
    res["refine_label_generation_ms"] = ev_time(lambda: mu.refine_label_generation_device(seg, heat, off, lab, gt, 10000, A))
    res["smooth_peak_pseudo_labels_ms"] = ev_time(
        lambda: mu.pseudo_label_generation_batch(gt, peak_extract_device(smoothing(heat), 15, 25), lab, 0.7, 6))
    return res


# ----------------------------------------------------------------------------- measured DRAM traffic (ncu) of the committed build
SWEEP_SOURCES = ["pamr_duo.cu", "pamr_lattice.cu", "pamr_lattice.cuh", "pamr_tma.cu", "pamr_sweep.cuh", "pamr_internal.cuh",
                 "tma.cuh", "common.cuh"]


def sweep_sources_sha():
    """Content hash of the files the sweep kernels are compiled from; profiles/traffic.json entries carry the hash of
    the build they were captured on (tools/record_traffic.py)."""
    h = hashlib.sha256()
    for f in SWEEP_SOURCES:
        with open(os.path.join(ROOT, "cl4wsis_b200", "csrc", f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload` from the tracked ncu --set full
    summary, or None (with the reason) when there is no capture of THIS build of the kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entries = json.load(f)["entries"]
    except Exception as e:  # noqa: BLE001
        return None, f"profiles/traffic.json unreadable: {e!r}"
    sha = sweep_sources_sha()
    stale = None
    for e in entries:
        if e["workload"] == workload and e["kernel"] == kernel:
            if e["sources_sha"] == sha:
                return float(e["dram_bytes_read"] + e["dram_bytes_write"]), f"{e['report']} (sources {sha})"
            stale = e
    if stale is not None:
        return None, f"capture {stale['report']} is of sources {stale['sources_sha']}, this build is {sha}"
    return None, "no ncu capture of this workload"


# ----------------------------------------------------------------------------- timing helpers
def ev_time(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def sweep_model(cfg, sweep_ms, peak):
    """Roofline figures of one propagation sweep: algorithmic bytes 4*H*W*(P + 2C)*B against the HBM
    peak, 2*C*P*H*W*B flops against the FP32 peak; `bound` by arithmetic intensity against 74.4 TF / peak."""
    B, C, H, W, P = cfg["B"], cfg["C"], cfg["H"], cfg["W"], 8 * len(cfg["dil"])
    nbytes = 4.0 * H * W * (P + 2 * C) * B
    flops = 2.0 * C * P * H * W * B
    t = sweep_ms * 1e-3
    return {"algorithmic_bytes_per_launch": nbytes, "flops_per_launch": flops, "achieved_gbs": nbytes / t / 1e9,
            "hbm_frac": nbytes / t / 1e9 / peak, "fp32_tflops": flops / t / 1e12, "fp32_frac_of_74.4": flops / t / 74.4e12,
            "intensity_flop_per_byte": flops / nbytes, "ridge_flop_per_byte": 74.4e3 / peak}


def run_workload(cl4, cfg, dev, first_image, steps, warmup, gen_device):
    """Short device-resident run of another BASELINE config (extra_workloads): images/s, ms per sweep, checksums."""
    B, C, H, W, T, dil = cfg["B"], cfg["C"], cfg["H"], cfg["W"], cfg["T"], cfg["dil"]
    img, mask, heat, off = (t.to(dev) for t in synth_inputs(cfg, first_image, device=gen_device))
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=T, dilations=dil, threshold=cfg["thr"], nms_kernel=cfg["nms"],
                               max_centers=max(256, 2 * cfg["Kc"]), device=dev)
    for _ in range(warmup):
        step.run(img, mask, heat, off)
    torch.cuda.synchronize()
    sweep_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in sweep_ev:
        a.record()
        b.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        step.run(img, mask, heat, off, sweep_events=sweep_ev[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sweep_ms = sum(a.elapsed_time(b) for a, b in sweep_ev) / (steps * T)
    refined, ids, counts, _ = step.run(img, mask, heat, off)
    torch.cuda.synchronize()
    out = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "mean_sweep_ms": sweep_ms,
           "sweep_kernel": step.sweep_kernel, "launches_per_step": step.launches_per_step,
           "checksums": {"mask": float(refined.double().sum()), "ids": float(ids.double().sum()),
                         "centres": int(counts.sum())}}
    del step, img, mask, heat, off, refined, ids, counts
    torch.cuda.empty_cache()
    return out


def run_sbd_pass(cdist, step, inputs, rank, world, dev, n_images=10582):
    """BASELINE configs[4]: one synthetic SBD-sized pass (10 582 images, C21, 512 x 512) of PAMR + centre NMS + grouping, images
    sharded contiguously over the ranks (cl4wsis_b200.dist.shard_bounds), batches of B, no data-path collective.  Each rank
    rolls its synthetic batch along the batch dimension from batch to batch; the tail batch runs full-size and only its real
    images are counted.  The timed region includes the rolls and the per-batch checksum reductions."""
    img, mask, heat, off = inputs
    B = img.shape[0]
    lo, hi = cdist.shard_bounds(n_images, rank, world)
    n_mine = hi - lo
    n_batches = (n_mine + B - 1) // B
    ck_mask = torch.zeros((), dtype=torch.float64, device=dev)
    ck_ids = torch.zeros((), dtype=torch.float64, device=dev)
    cdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n_batches):
        sft = k % B
        xi, mi, hi_, oi = (torch.roll(t, sft, 0) for t in (img, mask, heat, off)) if sft else (img, mask, heat, off)
        refined, ids, _, _ = step.run(xi, mi, hi_, oi)
        real = min(B, n_mine - k * B)
        ck_mask += refined[:real].sum(dtype=torch.float64)
        ck_ids += ids[:real].sum(dtype=torch.float64)
    e1.record()
    torch.cuda.synchronize()
    st = cdist.reduce_stats(n_mine, e0.elapsed_time(e1) / 1e3, float(ck_mask), float(ck_ids), device=dev)
    return {"workload": "sbd_pass_10582_c21_512", "images": st["images"], "pass_seconds": st["elapsed_s"],
            "value": st["images"] / st["elapsed_s"], "unit": UNIT, "n_gpus": world, "batches_per_rank": n_batches,
            "checksums": {"mask": st["checksum_mask"], "ids": st["checksum_ids"]},
            "note": "BASELINE configs[4]; timed region includes the batch rolls and the checksum reductions; max over ranks"}


def time_small_maps(cl4, dev):
    """The regime the trainer really runs PAMR in (train.py:372-385): softmax -> denorm + shrink -> PAMR(10,
    [1,2,4,8,12]) -> label gating -> pseudo_gtmask on feature-resolution maps.  Per shape: microseconds per
    phase1_pseudo_labels call (CUDA events over 50 calls, Python wrapper included), launches per call, and the fused
    PAMR kernel alone with its roofline (one-tile maps are launch/latency-bound; maps of several tiles are bound by the
    weight re-reads from L2: the numbers say so)."""
    from cl4wsis_b200.wss import single_stage as ss
    peak, _ = measured_peak_gbs()
    out = []
    for (B, C, h, w, Hi, Wi, note) in [(16, 21, 32, 32, 512, 512, "VOC, crop 512 / stride 16"),
                                       (16, 81, 56, 56, 448, 448, "coco-voc, 56x56 maps")]:
        g = torch.Generator(device=dev).manual_seed(7)
        images = torch.randn((B, 3, Hi, Wi), generator=g, device=dev)
        logits = 3 * torch.randn((B, C, h, w), generator=g, device=dev)
        l1h = (torch.rand((B, C - 1), generator=g, device=dev) < 0.1).float()
        mod = cl4.PAMR(10, [1, 2, 4, 8, 12]).to(dev)
        us = 1e3 * ev_time(lambda: ss.phase1_pseudo_labels(images, logits, l1h, mod), n=50)
        im = ss.denorm_resize(images, (h, w))
        soft = ss.softmax_channels(logits)
        us_pamr = 1e3 * ev_time(lambda: mod(im, soft), n=50)
        P, T = 40, 10
        nbytes = 4.0 * h * w * B * (3 + P + 2 * C)            # image + weights + masks in once, out once
        flops = 2.0 * C * P * T * h * w * B
        tiles = -(-h // 32) * -(-w // 32)
        extra = {}
        if tiles == 1:
            bound = ("launch + on-chip latency (all T sweeps inside one CTA per class group, weights in registers throughout; "
                     "HBM and FP32 fractions are both small)")
        else:
            # a CTA (cpb classes of one image: the rule of pamr_fused.cu pick_cpb, restated below) walks the map's tiles in every iteration and
            # re-reads a tile's weights (32 x 32 x P floats) from L2 at every tile switch
            pitch = 64 + 2 * 12                                   # two tiles + a frame of 12 (dilations <= 12)
            cpb_max = (227 * 1024) // (2 * pitch * pitch * 4)
            cpb = min(range(1, min(cpb_max, C) + 1), key=lambda k: (-(-(B * -(-C // k)) // 148) * (100 * k + 45), -k))
            l2 = 4.0 * 1024 * P * tiles * T * B * -(-C // cpb)
            extra = {"pamr_l2_weight_bytes": l2, "pamr_l2_weight_tbs": l2 / (us_pamr * 1e-6) / 1e12}
            bound = ("shared-memory pipe (l1tex 76 %, profiles/r03c_fused_ncu_summary.txt: one LDS.32 wavefront per 32 FMAs, no source "
                     "reuse in this mapping) on top of the weight re-reads from L2 -- a tile's weights per (tile, iteration, CTA): "
                     "pamr_l2_weight_bytes per call; HBM and FP32 fractions are small")
        out.append({"shape": f"B{B} C{C} {h}x{w} D5 T10 (images {Hi}x{Wi}; {note})", "phase1_pseudo_labels_us": us,
                    "launches_per_call": getattr(ss, "PHASE1_LAUNCHES", None), "pamr_call_us": us_pamr,
                    "pamr_launches": 2, "pamr_algorithmic_bytes": nbytes,
                    "pamr_hbm_frac": nbytes / (us_pamr * 1e-6) / 1e9 / peak,
                    "pamr_fp32_frac_of_74.4": flops / (us_pamr * 1e-6) / 74.4e12,
                    "bound": bound, **extra})
    return out


def time_torch_gpu(cfg, dev, n_images=4):
    """The same step restated with stock PyTorch ops (oracle/torch_ref.py = what the reference executes on CUDA tensors),
    fp32 convolutions (TF32 off), on `n_images` images of the workload; and PAMR alone in the trainer's 32x32 regime."""
    from oracle import torch_ref as tr
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        img, mask, heat, off = (t.to(dev) for t in synth_inputs(cfg, 0, n_images, device=dev))
        with torch.no_grad():
            ms = ev_time(lambda: tr.pseudo_label_step(img, mask, heat, off, cfg["T"], cfg["dil"], cfg["thr"], cfg["nms"]), n=3, warm=1)
            x = torch.rand((16, 3, 32, 32), device=dev)
            m = torch.rand((16, 21, 32, 32), device=dev).softmax(1)
            us_small = 1e3 * ev_time(lambda: tr.pamr(x, m, 10, [1, 2, 4, 8, 12]), n=20)
        del img, mask, heat, off
        torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return {"value": n_images / (ms * 1e-3), "unit": UNIT, "images": n_images, "ms_per_step": ms,
            "kind": "stock PyTorch ops (F.pad + F.conv2d + std + softmax; F.max_pool2d + nonzero; torch.norm + argmin), fp32, "
                    "oracle/torch_ref.py", "pamr_b16_c21_32x32_d5_us": us_small}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400, help="timed steps (default: ~2 s of device time at config 2)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="voc_b16_c21_512", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-images", type=int, default=64, help="images in the bounded cpu_baseline sample")
    ap.add_argument("--ref-images", type=int, default=0, help="images per step of the --impl reference arm (0: the whole batch "
                    "when the run then ends within ~3 minutes, else a bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callers", action="store_true", help="skip the timing of the widened rows (refine / pseudo labels)")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / small_map / extra_workloads / torch_gpu_baseline")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    _divert_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return

    import cl4wsis_b200 as cl4
    from cl4wsis_b200 import dist as cdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cdist.pin_to_local_cpus(local)       # cores / memory of the GPU's NUMA node (the e2e leg is host-memory bound)
    cdist.init_from_env("nccl")
    cl4._lib.load()

    B, C, H, W, T, dil = cfg["B"], cfg["C"], cfg["H"], cfg["W"], cfg["T"], cfg["dil"]
    P = 8 * len(dil)
    # rank r owns the images r*B .. r*B+B-1 of the job (seeded by global image index)
    h_img, h_mask, h_heat, h_off = (t.pin_memory() for t in synth_inputs(cfg, rank * B))
    img, mask, heat, off = (t.to(dev) for t in (h_img, h_mask, h_heat, h_off))
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=T, dilations=dil, threshold=cfg["thr"], nms_kernel=cfg["nms"],
                               max_centers=max(256, 2 * cfg["Kc"]), device=dev)

    kernel, launches_per_step = step.sweep_kernel, step.launches_per_step

    W_, K_ = max(args.warmup, 3), args.steps
    for _ in range(W_):
        step.run(img, mask, heat, off)
    torch.cuda.synchronize()

    # ---- device-resident timing: K steps, CUDA events, barrier + sync on both sides
    sweep_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_)]
    for a, b in sweep_ev:  # create the CUDA events now; the library re-records them around the sweeps
        a.record()
        b.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    cdist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0.record()
    for i in range(K_):
        step.run(img, mask, heat, off, sweep_events=sweep_ev[i])
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    cdist.barrier()
    elapsed = e0.elapsed_time(e1) / 1e3
    sweep_ms = sum(a.elapsed_time(b) for a, b in sweep_ev) / (K_ * T) if T > 0 else float("nan")

    refined, ids, counts, _ = step.run(img, mask, heat, off)
    torch.cuda.synchronize()
    ck_mask, ck_ids = float(refined.double().sum()), float(ids.double().sum())
    stats = cdist.reduce_stats(B * K_, elapsed, ck_mask, ck_ids, device=dev)
    value = stats["images"] / stats["elapsed_s"]

    # ---- sustained block: the driver's K may give a timed region of 0.1 s, too short for power / thermal steady state
    sustained = None
    if not args.no_extras and elapsed < 2.0:
        n_sus = int(math.ceil(2.2 / (elapsed / K_)))
        sclk = ClockSampler(local)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cdist.barrier()
        torch.cuda.synchronize()
        sclk.start()
        s0.record()
        for _ in range(n_sus):
            step.run(img, mask, heat, off)
        s1.record()
        torch.cuda.synchronize()
        sc = sclk.stop()
        sus = cdist.reduce_stats(B * n_sus, s0.elapsed_time(s1) / 1e3, 0.0, 0.0, device=dev)
        sustained = {"value": sus["images"] / sus["elapsed_s"], "unit": UNIT, "steps": n_sus, "seconds": sus["elapsed_s"],
                     "clocks": sc}

    # ---- end to end through host buffers (pinned H2D of inputs + D2H of results every step)
    e2e = None
    if not args.no_e2e:
        pipe = cl4.HostPseudoLabelPipeline(B, C, H, W, n_slots=3, num_iter=T, dilations=dil, threshold=cfg["thr"],
                                           nms_kernel=cfg["nms"], max_centers=max(256, 2 * cfg["Kc"]))
        n_e2e = K_ if K_ * 0.0105 >= 1.0 else int(math.ceil(1.0 / 0.0105))   # at least ~1 s of pipelined steps
        for _ in range(W_):
            pipe.submit(h_img, h_mask, h_heat, h_off)
        pipe.drain()
        torch.cuda.synchronize()
        cdist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            pipe.submit(h_img, h_mask, h_heat, h_off)
        pipe.drain()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        es = cdist.reduce_stats(B * n_e2e, dt, 0.0, 0.0, device=dev)
        e2e = {"value": es["images"] / es["elapsed_s"], "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "steps": n_e2e,
               "timing": "wall clock around the pipelined steps, sync on both sides, max over ranks"}
        # the floor of this leg: the same bytes per step moved both ways concurrently with no kernels in between
        fl = pipe.copy_floor(h_img, h_mask, h_heat, h_off, n=max(10, min(50, n_e2e)))
        fs = cdist.reduce_stats(B * fl["steps"], fl["seconds"], 0.0, 0.0, device=dev)
        e2e["copy_floor"] = {"value": fs["images"] / fs["elapsed_s"], "unit": UNIT,
                             "what": "H2D of the inputs and D2H of the results of a step on two streams, no kernels, all ranks at once",
                             "cpu_affinity": cdist.affinity_note()}
        e2e["frac_of_copy_floor"] = e2e["value"] / e2e["copy_floor"]["value"]
        del pipe
        torch.cuda.empty_cache()

    sbd = None
    if args.workload == "voc_b16_c21_512" and not args.no_extras:  # every rank takes part
        sbd = run_sbd_pass(cdist, step, (img, mask, heat, off), rank, world, dev)

    callers = small = extras = torch_gpu = None
    if rank == 0 and not args.no_callers:
        callers = time_callers(cfg, dev, 4321)
    if rank == 0 and world == 1 and not args.no_extras:
        small = time_small_maps(cl4, dev)
        del step, img, mask, heat, off, refined, ids, counts
        step = None
        torch.cuda.empty_cache()
        extras = {}
        peak_e, _ = measured_peak_gbs()
        for name in ("coco_b16_c81_512", "hires_b16_c21_1024"):
            if name == args.workload:
                continue
            r = run_workload(cl4, WORKLOADS[name], dev, 0, steps=10, warmup=3, gen_device=dev)
            r["config"] = bench_config(name, WORKLOADS[name])
            r["sweep"] = sweep_model(WORKLOADS[name], r["mean_sweep_ms"], peak_e)
            r["traffic"], r["traffic_source"] = measured_traffic(name, r["sweep_kernel"])
            extras[name] = r
        torch_gpu = time_torch_gpu(cfg, dev)

    # leave the process group cleanly before anything is printed (NCCL warns on stderr otherwise)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        cdist.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    peak, peak_src = measured_peak_gbs()
    sm = sweep_model(cfg, sweep_ms, peak)
    pamr_bytes_iter = 4.0 * H * W * (3 + P + T * (P + 2 * C))      # per image
    pamr_flops = 2.0 * C * P * T * H * W
    per_img_s = stats["elapsed_s"] / (B * K_)
    tiles = B * ((H + 31) // 32) * ((W + 31) // 32)
    # shared-memory wavefronts (128 B) one launch moves: LDS / STS of the compute warps + the TMA box, per
    # (tile, class pair) for the class-pair sweep: window 420, group A 76 LDS.64 + 8 STS.64 per thread (4 warps x 2
    # wavefronts), group B 32 + 24 + 4 LDS.128 per thread (4 warps x 4 wavefronts)
    if kernel == "pamr_sweep_duo_kernel":
        smem_bytes = tiles * ((C + 1) // 2) * 128.0 * (420 + 4 * 2 * 84 + 4 * 4 * 60)
    elif kernel == "pamr_sweep_lattice_kernel":
        smem_bytes = tiles * C * (128 * 84 * 4.0 + 128 * 60 * 8.0 + 80 * 84 * 4.0)
    elif len(dil) == 6:
        smem_bytes = tiles * C * (1024 * 143.0 + 80 * 80 * 4.0)
    else:
        smem_bytes = float("nan")
    sm_hz = (clk.get("sm_mhz") or 1965) * 1e6
    traffic, traffic_src = measured_traffic(args.workload, kernel)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": 1e3 * stats["elapsed_s"] / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, cfg),
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": sm["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": sm["hbm_frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": sm["algorithmic_bytes_per_launch"], "mean_launch_ms": sweep_ms,
                     "launches_timed": K_ * T,
                     "note": "mean over the T sweeps of a PAMR call: the first is the one-class lattice kernel writing pair "
                             "cells, the other T-1 the class-pair kernel" if kernel == "pamr_sweep_duo_kernel" else None},
        "roofline_fp32": {"bound": "fp32", "kernel": kernel, "achieved": sm["fp32_tflops"], "peak": 74.4, "unit": "TFLOP/s",
                          "frac": sm["fp32_frac_of_74.4"], "intensity_flop_per_byte": sm["intensity_flop_per_byte"],
                          "ridge_flop_per_byte": sm["ridge_flop_per_byte"],
                          "binding": "fp32" if sm["intensity_flop_per_byte"] > sm["ridge_flop_per_byte"] else "hbm"},
        "path_roofline": {"bytes_iter_frac_of_hbm": pamr_bytes_iter / per_img_s / 1e9 / peak,
                          "fp32_tflops": pamr_flops / per_img_s / 1e12, "fp32_frac_of_74.4": pamr_flops / per_img_s / 74.4e12,
                          # what actually binds the sweep: shared-memory wavefronts against 128 B/clk/SM
                          "sweep_smem_bytes_per_launch": smem_bytes,
                          "sweep_smem_frac_of_peak": smem_bytes / (sweep_ms * 1e-3) / (148 * 128 * sm_hz)},
        "e2e": e2e, "gpu_launches": K_ * launches_per_step, "clocks": clk,
        "checksums": {"mask": stats["checksum_mask"], "ids": stats["checksum_ids"]},
    }
    if sustained is not None:
        line["sustained"] = sustained
    if callers is not None:
        line["callers"] = callers
    if small is not None:
        line["small_map"] = small
    if extras is not None:
        line["extra_workloads"] = extras
    if sbd is not None:
        line["sbd_pass"] = sbd
    if torch_gpu is not None:
        line["torch_gpu_baseline"] = torch_gpu
    if world == 1 and not args.no_cpu_baseline:
        ips, dt, cores = time_oracle(cfg, args.cpu_images, 0, threads=host_threads())
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_images} images of workload {args.workload}, oracle/cl4_oracle.c "
                                          f"(C+OpenMP), {dt:.1f} s"}
    _emit(line)


if __name__ == "__main__":
    main()
