"""Batched, sync-free pseudo-label step: PAMR -> centre NMS -> offset grouping.

One *unit of work*  is one image: ``PAMR(img, mask)`` (reference
train.py:379 -> wss/modules.py:133), ``find_instance_center(heat)`` and
``group_pixels(centres, offsets)`` (train.py:492 -> modules/utils.py:565,604).  The drop-in
functions in ``cl4wsis_b200.wss`` / ``cl4wsis_b200.modules`` keep the reference's batch-1,
one-sync-per-call behaviour; this class runs the same kernels over a whole batch with the
centre counts left on the device, so a step has no host synchronisation at all.
"""
import torch

from . import _lib


class PseudoLabelStep:
    """Pre-allocated workspace + launches for a fixed (B, C, H, W) batch shape on one GPU.

    Capacity: ``cl4_center_nms`` stores the first ``max_centers`` centres of an image (``counts`` holds the true total) and
    the grouping uses the stored ones, so an image with more centres than ``max_centers`` is grouped against a truncated list
    -- unlike the reference's ``find_instance_center`` + ``group_pixels``.  Nothing is synchronised inside ``run``; check
    ``overflowed()`` (a device tensor, no sync) or ``assert_no_overflow()`` (one sync) and re-run such images through the
    drop-in functions of ``cl4wsis_b200.modules.utils``, or size ``max_centers`` for the worst case
    (``H * W / nms_kernel**2`` bounds the number of strict maxima; plateaux can exceed it)."""

    def __init__(self, B, C, H, W, K=3, num_iter=10, dilations=(1, 2, 4, 8, 12, 24), threshold=0.3, nms_kernel=41,
                 max_centers=256, ignore=True, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("PseudoLabelStep needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.C, self.H, self.W, self.K = B, C, H, W, K
        self.num_iter, self.dil = int(num_iter), [int(d) for d in dilations]
        self.threshold, self.nms_kernel = float(threshold), int(nms_kernel)
        self.max_centers, self.empty_mode = int(max_centers), 0 if ignore else 1
        dev = self.device
        lib = self.lib
        self._dil_arr = _lib.int_array(self.dil)
        self.pamr_bytes = lib.cl4_pamr_scratch_bytes(B, K, C, H, W, len(self.dil), self.num_iter)
        self.nms_bytes = lib.cl4_center_nms_scratch_bytes(B, H, W)
        self.pamr_scratch = torch.empty(max(self.pamr_bytes, 1), dtype=torch.uint8, device=dev)
        self.nms_scratch = torch.empty(max(self.nms_bytes, 1), dtype=torch.uint8, device=dev)
        self.refined = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        self.ids = torch.empty((B, H, W), dtype=torch.int64, device=dev)
        self.centers = torch.zeros((B, self.max_centers, 2), dtype=torch.int64, device=dev)
        self.counts = torch.zeros((B,), dtype=torch.int32, device=dev)
        # launches per step.  Lattice sweeps (class-default dilations, the replicate padding happens inside the kernels): image
        # pad, weights, num_iter sweeps (the first one the one-class kernel writing pair cells, the others the class-pair
        # kernel; CL4_SWEEP=lattice1: the one-class kernel throughout), 2 NMS, grouping.  4-pixel TMA sweep: image pad, weights,
        # mask pad, num_iter sweeps, num_iter-1 frame rewrites, 2 NMS, grouping.
        import os
        mode = os.environ.get("CL4_SWEEP")
        lattice = (self.dil in ([1, 2, 4, 8, 12, 24], [1, 2, 4, 8, 12]) and K <= 3 and W % 4 == 0 and H * W > 64 * 64
                   and mode in (None, "", "lattice", "lattice1"))
        self.sweep_kernel = ("pamr_sweep_duo_kernel" if lattice and mode != "lattice1" and self.num_iter >= 2
                             else "pamr_sweep_lattice_kernel" if lattice else "pamr_sweep_tma_kernel")
        self.launches_per_step = (1 + 1 + self.num_iter + 2 + 1 if lattice
                                  else 1 + 1 + 1 + self.num_iter + max(self.num_iter - 1, 0) + 2 + 1)

    def run(self, img, mask, heat, offsets, fg=None, stream=None, sweep_events=None):
        """All arguments are contiguous fp32 CUDA tensors: img [B,K,H,W], mask [B,C,H,W],
        heat [B,1,H,W], offsets [B,2,H,W]; fg optional uint8 [B,H,W].  Returns
        (refined [B,C,H,W], ids [B,H,W] int64, counts [B] int32, centres [B,max,2] int64) —
        views of this object's buffers, valid until the next ``run``.  ``sweep_events`` (a pair
        of timing events) brackets the ``num_iter`` propagation sweeps on the launch stream."""
        lib, B, C, H, W = self.lib, self.B, self.C, self.H, self.W
        ts = torch.cuda.current_stream(self.device) if stream is None else stream
        st = _lib.ctypes.c_void_p(ts.cuda_stream)
        ev0 = ev1 = None
        if sweep_events is not None:  # CUDA events recorded around the num_iter sweeps on this stream
            ev0, ev1 = (_lib.ctypes.c_void_p(e.cuda_event) for e in sweep_events)
        _lib.check(lib.cl4_pamr_forward_timed(_lib.ptr(img), _lib.ptr(mask), _lib.ptr(self.refined),
                                              _lib.ptr(self.pamr_scratch), self.pamr_bytes, B, self.K, C, H, W,
                                              self._dil_arr, len(self.dil), self.num_iter, st, ev0, ev1), "PAMR")
        _lib.check(lib.cl4_center_nms(_lib.ptr(heat), self.threshold, 0.0, self.nms_kernel, B, H, W,
                                      _lib.ptr(self.centers), _lib.ptr(self.counts), self.max_centers,
                                      _lib.ptr(self.nms_scratch), self.nms_bytes, st), "center_nms")
        _lib.check(lib.cl4_group_pixels(_lib.ptr(self.centers), _lib.ptr(self.counts), 0, self.max_centers,
                                        _lib.ptr(offsets), _lib.ptr(fg), _lib.ptr(self.ids), B, H, W,
                                        self.empty_mode, st), "group_pixels")
        return self.refined, self.ids, self.counts, self.centers

    def overflowed(self):
        """[B] bool device tensor: images of the last ``run`` whose centre list was truncated to ``max_centers``."""
        return self.counts > self.max_centers

    def assert_no_overflow(self):
        """One host sync: raises if any image of the last ``run`` had more than ``max_centers`` centres."""
        bad = torch.nonzero(self.overflowed()).flatten().tolist()
        if bad:
            raise OverflowError(f"PseudoLabelStep: images {bad} have more than max_centers={self.max_centers} centres; their "
                                "instance ids were computed against a truncated centre list")


class HostPseudoLabelPipeline:
    """End-to-end path for HOST buffers: pinned host -> device copies, the step, and the
    device -> pinned host copies of the refined masks, ids and centre counts, double-buffered
    over three streams so that PCIe traffic overlaps the kernels."""

    def __init__(self, B, C, H, W, n_slots=2, **kw):
        self.slots = []
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        K = kw.get("K", 3)
        for _ in range(n_slots):
            step = PseudoLabelStep(B, C, H, W, device=dev, **kw)
            slot = dict(
                step=step,
                img=torch.empty((B, K, H, W), dtype=torch.float32, device=dev),
                mask=torch.empty((B, C, H, W), dtype=torch.float32, device=dev),
                heat=torch.empty((B, 1, H, W), dtype=torch.float32, device=dev),
                off=torch.empty((B, 2, H, W), dtype=torch.float32, device=dev),
                h_refined=torch.empty((B, C, H, W), dtype=torch.float32).pin_memory(),
                h_ids=torch.empty((B, H, W), dtype=torch.int64).pin_memory(),
                h_counts=torch.empty((B,), dtype=torch.int32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_run=torch.cuda.Event(), ev_out=torch.cuda.Event(),
            )
            self.slots.append(slot)
        self.i = 0
        s = self.slots[0]
        self.h2d_bytes = sum(s[k].numel() * s[k].element_size() for k in ("img", "mask", "heat", "off"))
        self.d2h_bytes = sum(s[k].numel() * s[k].element_size() for k in ("h_refined", "h_ids", "h_counts"))
        self.launches_per_step = s["step"].launches_per_step

    def submit(self, h_img, h_mask, h_heat, h_off):
        """Enqueue one batch held in (pinned) host tensors; returns the slot whose ``h_*``
        outputs are valid after ``slot['ev_out'].synchronize()``."""
        s = self.slots[self.i % len(self.slots)]
        self.i += 1
        s["ev_out"].synchronize()  # the slot's previous results have left the device
        with torch.cuda.stream(self.s_in):
            s["img"].copy_(h_img, non_blocking=True)
            s["mask"].copy_(h_mask, non_blocking=True)
            s["heat"].copy_(h_heat, non_blocking=True)
            s["off"].copy_(h_off, non_blocking=True)
            s["ev_in"].record(self.s_in)
        self.s_run.wait_event(s["ev_in"])
        refined, ids, counts, _ = s["step"].run(s["img"], s["mask"], s["heat"], s["off"], stream=self.s_run)
        s["ev_run"].record(self.s_run)
        self.s_out.wait_event(s["ev_run"])
        with torch.cuda.stream(self.s_out):
            s["h_refined"].copy_(refined, non_blocking=True)
            s["h_ids"].copy_(ids, non_blocking=True)
            s["h_counts"].copy_(counts, non_blocking=True)
            s["ev_out"].record(self.s_out)
        return s

    def drain(self):
        for s in self.slots:
            s["ev_out"].synchronize()

    def copy_floor(self, h_img, h_mask, h_heat, h_off, n=20):
        """The floor of the end-to-end leg on this box: the copies of ``submit`` (inputs host -> device on one stream,
        results device -> host on another, concurrently) with no kernels in between.  -> {"steps", "seconds"}."""
        import time

        def one(s):
            with torch.cuda.stream(self.s_in):
                s["img"].copy_(h_img, non_blocking=True)
                s["mask"].copy_(h_mask, non_blocking=True)
                s["heat"].copy_(h_heat, non_blocking=True)
                s["off"].copy_(h_off, non_blocking=True)
            with torch.cuda.stream(self.s_out):
                s["h_refined"].copy_(s["step"].refined, non_blocking=True)
                s["h_ids"].copy_(s["step"].ids, non_blocking=True)
                s["h_counts"].copy_(s["step"].counts, non_blocking=True)

        self.drain()
        for i in range(2):
            one(self.slots[i % len(self.slots)])
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        for i in range(n):
            one(self.slots[i % len(self.slots)])
        torch.cuda.synchronize(self.device)
        return {"steps": n, "seconds": time.perf_counter() - t0}
