"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs the read-only checkout at /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these
fixtures are what pins the oracle (and, through it, the CUDA path).  Inputs are
stored alongside outputs, so the fixtures do not depend on any RNG stream.
Nothing at test/bench run time reads /root/reference.
"""
import importlib.util
import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

REF = os.environ.get("CL4WSIS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """Import the hot-path modules of the reference in place (SURVEY §8c workarounds)."""
    sys.path.insert(0, REF)
    mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    pp.get = None  # wss/modules.py:1 imports an unused name
    sys.modules.setdefault("matplotlib", mp)
    sys.modules.setdefault("matplotlib.pyplot", pp)
    import wss.modules as wm
    import wss.utils as wu
    spec = importlib.util.spec_from_file_location("ref_modules_utils", os.path.join(REF, "modules/utils.py"))
    mu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mu)
    return wm, wu, mu


def natural_image(rng, B, H, W):
    """8-bit-quantised, smooth-ish RGB in [0,1] (SURVEY §8d synthetic inputs)."""
    lo = torch.from_numpy(rng.integers(0, 256, (B, 3, max(2, H // 8), max(2, W // 8))).astype(np.float32))
    up = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
    return (up.round().clamp(0, 255) / 255.0).numpy().astype(np.float32)


def soft_mask(rng, B, C, H, W):
    lo = torch.from_numpy(rng.standard_normal((B, C, max(2, H // 8), max(2, W // 8))).astype(np.float32))
    up = torch.nn.functional.interpolate(3.0 * lo, size=(H, W), mode="bilinear", align_corners=False)
    return up.softmax(1).numpy().astype(np.float32)


def gaussian_heat(rng, H, W, n, sigma=6.0, amp=None):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    heat = np.zeros((H, W), np.float32)
    cs = []
    for i in range(n):
        cy, cx = int(rng.integers(0, H)), int(rng.integers(0, W))
        a = np.float32(amp[i]) if amp is not None else np.float32(rng.uniform(0.35, 1.0))
        g = a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sigma * sigma)).astype(np.float32)
        heat = np.maximum(heat, g)
        cs.append((cy, cx))
    return heat, np.array(cs, np.int64).reshape(-1, 2)


def offsets_to(rng, H, W, centres, noise=1.0):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    if len(centres) == 0:
        return rng.standard_normal((2, H, W)).astype(np.float32)
    d = (yy[None] - centres[:, 0, None, None]) ** 2 + (xx[None] - centres[:, 1, None, None]) ** 2
    near = d.argmin(0)
    oy = centres[near, 0] - yy + noise * rng.standard_normal((H, W))
    ox = centres[near, 1] - xx + noise * rng.standard_normal((H, W))
    return np.stack([oy, ox]).astype(np.float32)


def main():
    wm, wu, mu = load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    rng = np.random.default_rng(20261018)
    out = {}

    # ---------------- PAMR (wss/modules.py:122-152) ----------------
    pamr_cases = {
        "pamr_d6":        dict(B=2, C=5, H=40, W=36, dil=[1, 2, 4, 8, 12, 24], T=10, img="natural"),
        "pamr_d5":        dict(B=1, C=3, H=33, W=47, dil=[1, 2, 4, 8, 12], T=10, img="rand"),
        "pamr_tiny_d6":   dict(B=2, C=4, H=16, W=12, dil=[1, 2, 4, 8, 12, 24], T=10, img="natural"),  # d > H
        "pamr_1iter":     dict(B=1, C=2, H=24, W=24, dil=[1, 2], T=1, img="rand"),
        "pamr_flat":      dict(B=1, C=3, H=20, W=20, dil=[1, 2, 4], T=3, img="flat"),
        "pamr_resize":    dict(B=1, C=4, H=32, W=28, dil=[1, 2, 4, 8, 12], T=10, img="natural", mh=9, mw=11),
        "pamr_c21_64":    dict(B=1, C=21, H=64, W=64, dil=[1, 2, 4, 8, 12, 24], T=10, img="natural"),
    }
    for name, c in pamr_cases.items():
        B, C, H, W = c["B"], c["C"], c["H"], c["W"]
        if c["img"] == "natural":
            x = natural_image(rng, B, H, W)
        elif c["img"] == "rand":
            x = rng.random((B, 3, H, W)).astype(np.float32)
        else:  # flat patches: zero std on part of the image
            x = np.full((B, 3, H, W), 0.5, np.float32)
            x[:, :, :, W // 2:] = rng.random((B, 3, H, W - W // 2)).astype(np.float32)
        m = soft_mask(rng, B, C, c.get("mh", H), c.get("mw", W))
        mod = wm.PAMR(num_iter=c["T"], dilations=c["dil"])
        with torch.no_grad():
            y = mod(torch.from_numpy(x), torch.from_numpy(m)).numpy()
        out[name + "__x"], out[name + "__mask"], out[name + "__out"] = x, m, y
        out[name + "__dil"], out[name + "__T"] = np.array(c["dil"], np.int32), np.array(c["T"], np.int32)
        print(name, y.shape, float(y.sum(1).mean()))

    # affinity weights alone (exposes :141-145 without the iterations)
    x = natural_image(rng, 1, 24, 30)
    mod = wm.PAMR(num_iter=1, dilations=[1, 2, 4, 8, 12, 24])
    with torch.no_grad():
        xt = torch.from_numpy(x)
        w = torch.softmax((-mod.aff_x(xt) / (1e-8 + 0.1 * mod.aff_std(xt))).mean(1, keepdim=True), 2)[:, 0].numpy()
    out["weights_d6__x"], out["weights_d6__w"] = x, w

    # ---------------- peak_extract (wss/utils.py:3-25) ----------------
    pk_cases = {
        "peak_k15": dict(B=2, C=3, H=64, W=48, n=6, kernel=15, K=25),
        "peak_k5":  dict(B=1, C=2, H=32, W=32, n=4, kernel=5, K=5),
        "peak_k3_neg": dict(B=1, C=2, H=20, W=24, n=0, kernel=3, K=7),
    }
    for name, c in pk_cases.items():
        heat = np.zeros((c["B"], c["C"], c["H"], c["W"]), np.float32)
        for b in range(c["B"]):
            for ch in range(c["C"]):
                if c["n"]:
                    amp = np.linspace(0.95, 0.4, c["n"]) + rng.uniform(0, 0.01, c["n"])
                    heat[b, ch], _ = gaussian_heat(rng, c["H"], c["W"], c["n"], sigma=3.0, amp=amp)
                else:  # signed noise: negatives and non-kept zeros compete in top-k
                    heat[b, ch] = rng.standard_normal((c["H"], c["W"])).astype(np.float32)
        sc, ys, xs = wu.peak_extract(torch.from_numpy(heat), kernel=c["kernel"], K=c["K"])
        out[name + "__heat"], out[name + "__scores"], out[name + "__ys"], out[name + "__xs"] = heat, sc, ys, xs
        out[name + "__kernel"], out[name + "__K"] = np.array(c["kernel"], np.int32), np.array(c["K"], np.int32)
        print(name, sc.shape, sc.dtype, ys.dtype)
    # SURVEY §8c ⑧
    heat = np.zeros((1, 1, 32, 32), np.float32)
    heat[0, 0, 5, 7], heat[0, 0, 20, 3], heat[0, 0, 20, 25] = 0.9, 0.8, 0.8
    sc, ys, xs = wu.peak_extract(torch.from_numpy(heat), kernel=5, K=5)
    out["peak_kat8__heat"], out["peak_kat8__scores"], out["peak_kat8__ys"], out["peak_kat8__xs"] = heat, sc, ys, xs
    out["peak_kat8__kernel"], out["peak_kat8__K"] = np.array(5, np.int32), np.array(5, np.int32)

    # ---------------- find_instance_center (modules/utils.py:463-502) ----------------
    kat = np.zeros((1, 1, 64, 64), np.float32)
    kat[0, 0, 10, 10], kat[0, 0, 10, 12], kat[0, 0, 40, 40], kat[0, 0, 41, 41], kat[0, 0, 5, 60] = .9, .8, .5, .5, .05
    i = 0
    for k in (3, 5, 41):
        out[f"center_{i}__heat"] = kat
        out[f"center_{i}__args"] = np.array([0.3, k, -1], np.float64)
        out[f"center_{i}__ctr"] = mu.find_instance_center(torch.from_numpy(kat.copy()), 0.3, k, None).numpy()
        i += 1
    for (H, W, n, thr, k, topk) in [(48, 56, 5, 0.3, 41, None), (48, 56, 5, 0.1, 7, 10000), (37, 29, 12, 0.3, 3, None),
                                    (64, 64, 0, 0.3, 5, None), (50, 70, 30, 0.1, 3, None), (96, 96, 8, 0.3, 41, 10000)]:
        heat, _ = gaussian_heat(rng, H, W, n) if n else (np.zeros((H, W), np.float32), None)
        heat = (heat + 0.02 * rng.random((H, W)).astype(np.float32))[None, None]
        out[f"center_{i}__heat"] = heat
        out[f"center_{i}__args"] = np.array([thr, k, -1 if topk is None else topk], np.float64)
        out[f"center_{i}__ctr"] = mu.find_instance_center(torch.from_numpy(heat.copy()), thr, k, topk).numpy()
        i += 1
    # plateau + quantised heat (many exact ties)
    heat = (np.round(gaussian_heat(rng, 40, 40, 6, sigma=4.0)[0] * 8) / 8).astype(np.float32)[None, None]
    out[f"center_{i}__heat"], out[f"center_{i}__args"] = heat, np.array([0.2, 5, -1], np.float64)
    out[f"center_{i}__ctr"] = mu.find_instance_center(torch.from_numpy(heat.copy()), 0.2, 5, None).numpy()
    i += 1
    # SURVEY §8c ⑦: degenerate top_k branch (prints the centre count)
    heat = np.zeros((1, 1, 64, 64), np.float32)
    heat[0, 0, 2::6, 2::6] = (0.5 + 0.4 * rng.random((11, 11))).astype(np.float32)
    with redirect_stdout(io.StringIO()) as so:
        c = mu.find_instance_center(torch.from_numpy(heat.copy()), 0.3, 3, 20).numpy()
    out[f"center_{i}__heat"], out[f"center_{i}__args"], out[f"center_{i}__ctr"] = heat, np.array([0.3, 3, 20], np.float64), c
    out[f"center_{i}__printed"] = np.array(int(so.getvalue().strip()), np.int64)
    i += 1
    out["center__n"] = np.array(i, np.int32)

    # ---------------- group_pixels (modules/utils.py:505-542) ----------------
    j = 0
    out[f"group_{j}__ctr"] = np.array([[1, 2], [1, 4]], np.int64)
    out[f"group_{j}__off"] = np.zeros((1, 2, 4, 8), np.float32)
    j += 1
    for (H, W, Kc, noise) in [(48, 56, 5, 1.0), (64, 64, 1, 3.0), (40, 72, 50, 2.0), (33, 31, 200, 4.0), (64, 64, 7, 0.0)]:
        ctr = np.stack([rng.integers(0, H, Kc), rng.integers(0, W, Kc)], 1).astype(np.int64)
        off = offsets_to(rng, H, W, ctr, noise)[None]
        if noise == 0.0:  # integer offsets: many exact distance ties
            off = np.round(off)
            ctr[1] = ctr[0]  # duplicate centre: lowest index must win
        out[f"group_{j}__ctr"], out[f"group_{j}__off"] = ctr, off.astype(np.float32)
        j += 1
    for q in range(j):
        out[f"group_{q}__ids"] = mu.group_pixels(torch.from_numpy(out[f"group_{q}__ctr"]),
                                                 torch.from_numpy(out[f"group_{q}__off"])).numpy()
    out["group__n"] = np.array(j, np.int32)

    # ---------------- get_instance_segmentation (modules/utils.py:545-606) ----------------
    g = 0
    for (H, W, n, thr, k, ignore, beta, empty) in [(64, 72, 4, 0.3, 41, True, 0, False), (64, 72, 0, 0.3, 41, True, 0, True),
                                                   (64, 72, 0, 0.3, 41, False, 0, True), (80, 80, 3, 0.3, 41, True, 3.0, False),
                                                   (80, 80, 6, 0.1, 5, True, 5, False)]:
        heat, cs = gaussian_heat(rng, H, W, max(n, 1))
        if empty:
            heat *= 0.1
        off = offsets_to(rng, H, W, cs, 0.5)[None]
        fg = (rng.random((1, H, W)) > 0.3)
        hm = heat[None, None].copy()
        ids = mu.get_instance_segmentation(torch.from_numpy(fg), torch.from_numpy(hm), torch.from_numpy(off),
                                           threshold=thr, nms_kernel=k, top_k=None, ignore=ignore, beta=beta).numpy()
        out[f"inst_{g}__fg"], out[f"inst_{g}__heat"], out[f"inst_{g}__off"] = fg, heat[None, None], off
        out[f"inst_{g}__heat_after"] = hm  # the reference marks merged cluster centres in place
        out[f"inst_{g}__args"] = np.array([thr, k, float(ignore), beta], np.float64)
        out[f"inst_{g}__ids"] = ids
        g += 1
    # centre clustering exercised for real (modules/utils.py:567-594): exact offsets give a 21-pixel
    # weak-offset disk around every planted centre; heat peaks exist only for some of them
    def cluster_case(H, W, centres, peaks, base, thr, k, beta, fg_holes):
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        cs = np.array(centres, np.int64)
        off = offsets_to(rng, H, W, cs, 0.0)[None]
        heat = np.full((H, W), base, np.float32)
        for (cy, cx, a) in peaks:
            heat = np.maximum(heat, a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72.0).astype(np.float32))
        fg = np.ones((1, H, W), bool)
        for (y0, y1, x0, x1) in fg_holes:
            fg[0, y0:y1, x0:x1] = False
        return fg, heat[None, None], off, thr, k, True, beta

    extra = [
        # two NMS centres + two far cluster centres (one added, one within 100 px of an NMS centre)
        cluster_case(200, 260, [(30, 40), (40, 230), (170, 60), (60, 90)], [(30, 40, 0.9), (40, 230, 0.8)], 0.06, 0.3, 41, 3.0, []),
        # no NMS centre at all: the cluster centres become the centres
        cluster_case(120, 150, [(20, 30), (90, 120)], [], 0.1, 0.3, 41, 5, []),
        # cluster centre heat below 0.05 is ignored; a hole in fg splits a disk
        cluster_case(180, 180, [(30, 30), (150, 150)], [(30, 30, 0.9)], 0.01, 0.3, 5, 5, [(148, 153, 150, 151)]),
        # beta too small for the 21-pixel disks to pass (21-1 < area < 21+1 still passes 21)
        cluster_case(160, 200, [(25, 25), (140, 170)], [(25, 25, 0.7)], 0.2, 0.3, 41, 1, []),
    ]
    for (fg, hm0, off, thr, k, ignore, beta) in extra:
        hm = hm0.copy()
        ids = mu.get_instance_segmentation(torch.from_numpy(fg), torch.from_numpy(hm), torch.from_numpy(off),
                                           threshold=thr, nms_kernel=k, top_k=None, ignore=ignore, beta=beta).numpy()
        out[f"inst_{g}__fg"], out[f"inst_{g}__heat"], out[f"inst_{g}__off"] = fg, hm0, off
        out[f"inst_{g}__heat_after"] = hm
        out[f"inst_{g}__args"] = np.array([thr, k, float(ignore), beta], np.float64)
        out[f"inst_{g}__ids"] = ids
        print("cluster case", g, "ids max", int(ids.max()), "cells marked", int((hm != hm0).sum()))
        g += 1
    out["inst__n"] = np.array(g, np.int32)

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays; torch", torch.__version__)


if __name__ == "__main__":
    main()
