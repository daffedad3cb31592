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

`--impl reference` times the oracle port alone (the reference itself is Python and does not
travel to the GPU box).
"""
import argparse
import json
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
def synth_inputs(cfg, seed, n_images=None):
    """SURVEY §8d synthetic inputs on the CPU (seed = 1234 + rank): natural-like 8-bit image,
    dense softmax masks (what the trainer feeds PAMR: train.py:373-379), gaussian centre
    heat-map with Kc planted centres (sigma 6), offsets to the nearest centre + N(0,1)."""
    B = cfg["B"] if n_images is None else n_images
    C, H, W, Kc = cfg["C"], cfg["H"], cfg["W"], cfg["Kc"]
    g = torch.Generator().manual_seed(seed)
    lo = torch.randint(0, 256, (B, 3, H // 8, W // 8), generator=g).float()
    img = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False).round().clamp(0, 255) / 255.0
    mlo = torch.randn((B, C, H // 8, W // 8), generator=g)
    mask = torch.nn.functional.interpolate(3.0 * mlo, size=(H, W), mode="bilinear", align_corners=False).softmax(1)
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    heat = torch.zeros((B, 1, H, W))
    off = torch.empty((B, 2, H, W))
    for b in range(B):
        cy = torch.randint(0, H, (Kc,), generator=g).float().view(Kc, 1, 1)
        cx = torch.randint(0, W, (Kc,), generator=g).float().view(Kc, 1, 1)
        amp = (0.5 + 0.5 * torch.rand((Kc,), generator=g)).view(Kc, 1, 1)
        d2 = (yy - cy) ** 2 + (xx - cx) ** 2                      # [Kc,H,W]
        heat[b, 0] = (amp * torch.exp(-d2 / (2 * 6.0 * 6.0))).amax(0)
        near = d2.argmin(0)
        off[b, 0] = cy.view(-1)[near] - yy.expand(1, H, W)[0] + torch.randn((H, W), generator=g)
        off[b, 1] = cx.view(-1)[near] - xx.expand(1, H, W)[0] + torch.randn((H, W), generator=g)
        del d2
    return img.contiguous(), mask.contiguous(), heat.contiguous(), off.contiguous()


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
def time_oracle(cfg, n_images, seed, repeats=1, threads=None):
    """The oracle port on host cores over `n_images` images of the workload -> images/s."""
    import oracle as orc
    orc.build()
    if threads:
        orc.set_num_threads(threads)
    img, mask, heat, off = (t.numpy() for t in synth_inputs(cfg, seed, n_images))
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
    if rank != 0:
        return
    n_img = args.ref_images
    for _ in range(args.warmup):
        time_oracle(cfg, 1, 999, threads=host_threads())
    t_total, cores = 0.0, None
    for i in range(args.steps):
        ips, dt, cores = time_oracle(cfg, n_img, 1234 + i, threads=host_threads())
        t_total += dt
    value = args.steps * n_img / t_total
    sample = f"{n_img} images of workload {args.workload} per step x {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **{k: cfg[k] for k in ("B", "C", "H", "W", "T")}, "dilations": cfg["dil"],
                   "nms_kernel": cfg["nms"], "threshold": cfg["thr"], "images_per_step": n_img},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/PyTorch and cannot travel to the GPU box; this is oracle/cl4_oracle.c "
                "(C + OpenMP port, ~12x faster than the reference's own torch-CPU path measured in SURVEY §6)",
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


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="voc_b16_c21_512", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-images", type=int, default=64, help="images in the bounded cpu_baseline sample")
    ap.add_argument("--ref-images", type=int, default=4, help="images per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callers", action="store_true", help="skip the timing of the widened rows (refine / pseudo labels)")
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
    cdist.init_from_env("nccl")
    cl4._lib.load()

    B, C, H, W, T, dil = cfg["B"], cfg["C"], cfg["H"], cfg["W"], cfg["T"], cfg["dil"]
    P = 8 * len(dil)
    h_img, h_mask, h_heat, h_off = (t.pin_memory() for t in synth_inputs(cfg, 1234 + rank))
    img, mask, heat, off = (t.to(dev) for t in (h_img, h_mask, h_heat, h_off))
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=T, dilations=dil, threshold=cfg["thr"], nms_kernel=cfg["nms"],
                               max_centers=max(256, 2 * cfg["Kc"]), device=dev)

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

    # ---- end to end through host buffers (pinned H2D of inputs + D2H of results every step)
    e2e = None
    if not args.no_e2e:
        pipe = cl4.HostPseudoLabelPipeline(B, C, H, W, n_slots=3, num_iter=T, dilations=dil, threshold=cfg["thr"],
                                           nms_kernel=cfg["nms"], max_centers=max(256, 2 * cfg["Kc"]))
        for _ in range(W_):
            pipe.submit(h_img, h_mask, h_heat, h_off)
        pipe.drain()
        torch.cuda.synchronize()
        cdist.barrier()
        t0 = time.perf_counter()
        for _ in range(K_):
            pipe.submit(h_img, h_mask, h_heat, h_off)
        pipe.drain()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        es = cdist.reduce_stats(B * K_, dt, 0.0, 0.0, device=dev)
        e2e = {"value": es["images"] / es["elapsed_s"], "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "timing": "wall clock around K pipelined steps, sync on both sides"}
        del pipe

    callers = None
    if rank == 0 and not args.no_callers:
        callers = time_callers(cfg, dev, 4321)

    # leave the process group cleanly before anything is printed (NCCL warns on stderr otherwise)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        cdist.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    peak, peak_src = measured_peak_gbs()
    sweep_bytes = 4.0 * H * W * (P + 2 * C) * B          # algorithmic bytes of one sweep launch
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    pamr_bytes_iter = 4.0 * H * W * (3 + P + T * (P + 2 * C))      # per image
    pamr_flops = 2.0 * C * P * T * H * W
    per_img_s = stats["elapsed_s"] / (B * K_)
    tiles = B * ((H + 31) // 32) * ((W + 31) // 32)
    # shared-memory bytes one launch moves: LDS/STS of the compute warps + the TMA box per (tile, class)
    lattice = dil == [1, 2, 4, 8, 12, 24] and os.environ.get("CL4_SWEEP") in (None, "", "lattice")
    if lattice:    # group A 128 thr x (76 + 8) x 4 B, group B 128 thr x (56 + 4) x 8 B, window 80 x 84 fp32
        smem_bytes = tiles * C * (128 * 84 * 4.0 + 128 * 60 * 8.0 + 80 * 84 * 4.0)
    elif len(dil) == 6:
        smem_bytes = tiles * C * (1024 * 143.0 + 80 * 80 * 4.0)
    else:
        smem_bytes = float("nan")
    sm_hz = (clk.get("sm_mhz") or 1965) * 1e6
    # dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch from the committed ncu --set full
    # capture of this workload (profiles/r01f_sweep_ncu_raw.csv: 1.557 GB read + 0.342 GB written by the lattice sweep;
    # profiles/r01b_sweep_ncu_raw.csv: 1.892 GB for the 4-pixel sweep); other workloads were not captured
    traffic = (1.900e9 if lattice else 1.892e9) if args.workload == "voc_b16_c21_512" else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": 1e3 * stats["elapsed_s"] / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "B_per_gpu": B, "C": C, "H": H, "W": W, "dilations": dil, "num_iter": T,
                   "nms_kernel": cfg["nms"], "threshold": cfg["thr"], "centres_per_image": cfg["Kc"],
                   "mask": "dense softmax over all classes", "l2": "inputs + scratch per step exceed L2 (no flush needed)"},
        "roofline": {"bound": "hbm", "kernel": "pamr_sweep_lattice" if lattice else "pamr_sweep_tma", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": sweep_bytes, "mean_launch_ms": sweep_ms},
        "path_roofline": {"bytes_iter_frac_of_hbm": pamr_bytes_iter / per_img_s / 1e9 / peak,
                          "fp32_tflops": pamr_flops / per_img_s / 1e12, "fp32_frac_of_74.4": pamr_flops / per_img_s / 74.4e12,
                          # what actually binds the sweep: 143 LDS words per 4 pixel-classes
                          # plus the 80x80 TMA window per 1024 pixel-classes, against 128 B/clk/SM
                          "sweep_smem_bytes_per_launch": smem_bytes,
                          "sweep_smem_frac_of_peak": smem_bytes / (sweep_ms * 1e-3) / (148 * 128 * sm_hz)},
        "e2e": e2e, "gpu_launches": K_ * step.launches_per_step, "clocks": clk,
        "checksums": {"mask": stats["checksum_mask"], "ids": stats["checksum_ids"]},
    }
    if callers is not None:
        line["callers"] = callers
    if world == 1 and not args.no_cpu_baseline:
        ips, dt, cores = time_oracle(cfg, args.cpu_images, 1234, threads=host_threads())
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_images} images of workload {args.workload}, oracle/cl4_oracle.c "
                                          f"(C+OpenMP), {dt:.1f} s"}
    _emit(line)


if __name__ == "__main__":
    main()
