"""Comparison of two peak_extract results computed from heat maps that agree only to fp32 rounding.

`keep = (max_pool(heat) == heat)` (reference wss/utils.py:8-11) is an exact float comparison, so on a bilinearly up-sampled
map (smooth ramps, clamped borders) pixels whose value ties with the window maximum to the last bit are peaks or not depending on
one rounding -- ATen's own CPU and CUDA up-sampling kernels disagree on them.  Two results are EQUIVALENT when
  * every entry present in both lists (same pixel) has the same score to `rtol`, and
  * every entry present in only one list is such a marginal pixel: in the oracle's up-sampled map its value is within `ulps`
    units in the last place of its k x k window maximum, or it lies below the other list's last score (pushed out of the
    top-K by marginal entries above it).
"""
import numpy as np


def window_max(a, k):
    r = (k - 1) // 2
    H, W = a.shape
    p = np.full((H + 2 * r, W + 2 * r), -np.inf, a.dtype)
    p[r:r + H, r:r + W] = a
    cols = np.max(np.stack([p[i:i + H] for i in range(k)]), 0)
    return np.max(np.stack([cols[:, j:j + W] for j in range(k)]), 0)


def assert_peaks_equivalent(got, want, up, kernel, rtol=1e-5, ulps=8, min_common=0.8):
    """got / want: (scores, ys, xs) arrays [B,C,K]; up: the oracle's up-sampled heat [B,C,H,W]."""
    gs, gy, gx = got
    ws, wy, wx = want
    B, C, K = ws.shape
    common = total = 0
    for b in range(B):
        for c in range(C):
            hm = window_max(up[b, c], kernel)
            marginal = np.abs(hm - up[b, c]) <= ulps * np.spacing(np.abs(hm).astype(np.float32))
            G = {(int(y), int(x)): float(s) for s, y, x in zip(gs[b, c], gy[b, c], gx[b, c]) if s > 0}
            Wd = {(int(y), int(x)): float(s) for s, y, x in zip(ws[b, c], wy[b, c], wx[b, c]) if s > 0}
            for p in G.keys() & Wd.keys():
                assert abs(G[p] - Wd[p]) <= rtol * abs(Wd[p]) + 1e-7, (b, c, p, G[p], Wd[p])
            lo_g = min(G.values()) if len(G) == K else 0.0
            lo_w = min(Wd.values()) if len(Wd) == K else 0.0
            for p in G.keys() - Wd.keys():
                assert marginal[p] or G[p] <= lo_w * (1 + rtol) + 1e-7, ("only ours", b, c, p, G[p])
            for p in Wd.keys() - G.keys():
                assert marginal[p] or Wd[p] <= lo_g * (1 + rtol) + 1e-7, ("only reference", b, c, p, Wd[p])
            common += len(G.keys() & Wd.keys())
            total += max(len(G), len(Wd))
    assert total == 0 or common >= min_common * total, (common, total)
