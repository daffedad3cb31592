"""Oracle vs the reference run live — only where /root/reference exists (the build
container).  Skipped on the GPU box; the committed fixtures cover it there."""
import os

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "wss")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    wm, wu, mu = mg.load_reference()
    if REF in sys.path:
        sys.path.remove(REF)  # keep the reference's top-level packages out of later imports
    return wm, wu, mu, mg


def test_group_pixels_norm_formula_many_shapes(ref, oracle):
    """ATen's CPU 2-vector norm == sqrt(fma(dx,dx,rn(dy*dy))), first-min argmin."""
    import torch
    _, _, mu, _ = ref
    rng = np.random.default_rng(7)
    total = 0
    for (H, W, Kc) in [(64, 64, 3), (31, 77, 16), (128, 96, 40), (17, 19, 200), (96, 128, 9), (256, 256, 24)]:
        ctr = np.stack([rng.integers(0, H, Kc), rng.integers(0, W, Kc)], 1).astype(np.int64)
        off = (rng.standard_normal((1, 2, H, W)) * 20).astype(np.float32)
        want = mu.group_pixels(torch.from_numpy(ctr), torch.from_numpy(off)).numpy()
        got = oracle.group_pixels(ctr, off)
        assert np.array_equal(got, want), (H, W, Kc, int((got != want).sum()))
        total += H * W * Kc
    assert total > 2_000_000


def test_pamr_random_shapes(ref, oracle):
    import torch
    wm, _, _, mg = ref
    rng = np.random.default_rng(11)
    for (B, C, H, W, dil) in [(1, 2, 25, 31, [1, 2, 4, 8, 12, 24]), (2, 3, 48, 40, [1, 2, 4, 8, 12]), (1, 1, 9, 7, [1, 3])]:
        x = mg.natural_image(rng, B, H, W)
        m = mg.soft_mask(rng, B, C, H, W)
        with torch.no_grad():
            want = wm.PAMR(10, dil)(torch.from_numpy(x), torch.from_numpy(m)).numpy()
        np.testing.assert_allclose(oracle.pamr(x, m, 10, dil), want, rtol=1e-5, atol=1e-7)


def test_center_nms_random(ref, oracle):
    import torch
    _, _, mu, mg = ref
    rng = np.random.default_rng(13)
    for (H, W, n, thr, k) in [(64, 80, 9, 0.3, 41), (40, 40, 20, 0.1, 3), (33, 65, 4, 0.5, 7), (16, 16, 2, 0.3, 41)]:
        heat, _ = mg.gaussian_heat(rng, H, W, n)
        heat = np.round(heat * 64) / 64  # exact ties
        want = mu.find_instance_center(torch.from_numpy(heat[None, None].copy()), thr, k, None).numpy()
        assert np.array_equal(oracle.find_instance_center(heat[None, None], thr, k), want)
