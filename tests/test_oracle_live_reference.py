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


def test_phase1_pieces_random_shapes(ref):
    """oracle/phase1.py against the reference's own `denorm` (utils/utils.py:26-41) and `pseudo_gtmask`
    (wss/single_stage.py:18-40) function bodies, and against torch's interpolate / softmax, on random shapes incl.
    all-equal planes (max ties), NaN-free extremes and every cutoff branch."""
    import importlib.util
    import torch
    import torch.nn.functional as F
    from oracle import phase1 as p1
    spec = importlib.util.spec_from_file_location(
        "make_golden_more", os.path.join(os.path.dirname(__file__), "golden", "make_golden_more.py"))
    mm = importlib.util.module_from_spec(spec)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    try:
        spec.loader.exec_module(mm)
    finally:
        sys.path.remove(os.path.join(os.path.dirname(__file__), "golden"))
        if REF in sys.path:
            sys.path.remove(REF)
    denorm = mm._ref_function("utils/utils.py", "denorm")
    pseudo_gtmask = mm._ref_function("wss/single_stage.py", "pseudo_gtmask")
    rng = np.random.default_rng(17)
    for (B, C, Hi, Wi, h, w) in [(2, 4, 40, 56, 10, 14), (1, 21, 64, 64, 32, 32), (3, 2, 9, 11, 9, 11), (1, 5, 33, 17, 2, 3)]:
        img = rng.standard_normal((B, 3, Hi, Wi)).astype(np.float32)
        assert np.array_equal(p1.denorm(img), denorm(torch.from_numpy(img)).numpy())
        assert np.array_equal(p1.denorm(img[0]), denorm(torch.from_numpy(img[0])).numpy())      # the [3,H,W] branch
        raw = p1.denorm(img)
        want = F.interpolate(torch.from_numpy(raw), (h, w), mode="bilinear", align_corners=True).numpy()
        np.testing.assert_allclose(p1.resize_bilinear_ac(raw, (h, w)), want, rtol=2e-6, atol=1e-6)
        logits = (4 * rng.standard_normal((B, C, h, w))).astype(np.float32)
        soft = torch.from_numpy(logits).softmax(1).numpy()
        np.testing.assert_allclose(p1.softmax_channels(logits), soft, rtol=2e-6, atol=1e-9)
        l1h = (rng.random((B, C - 1)) < 0.5).astype(np.float32)
        gated = p1.gate_labels(soft, l1h)
        g2 = torch.from_numpy(soft.copy())
        g2[:, 1:] *= torch.from_numpy(l1h)[:, :, None, None]
        assert np.array_equal(gated, g2.numpy())
        gated[0, -1] = 0.25                      # a constant plane: nothing exceeds its own scaled maximum
        for amb, top, bkg, low in [(True, 0.6, 0.7, 0.2), (False, 0.6, 0.6, 0.2), (True, 0.9, 0.1, 0.0), (True, 0.3, 0.3, 0.9)]:
            want = pseudo_gtmask(torch.from_numpy(gated.copy()), ambiguous=amb, cutoff_top=top, cutoff_bkg=bkg, cutoff_low=low).numpy()
            assert np.array_equal(p1.pseudo_gtmask(gated, amb, top, bkg, low), want), (B, C, amb, top, bkg, low)


def test_torch_ref_matches_live_reference(ref):
    """oracle/torch_ref.py (the stock-PyTorch restatement that travels to the GPU box as second oracle and second
    baseline) against the reference's own modules on CPU tensors: PAMR bit-exact (same ATen ops in the same order),
    centre lists, instance ids, peaks and smoothing exact."""
    import torch
    from oracle import torch_ref as tr
    wm, wu, mu, mg = ref
    rng = np.random.default_rng(23)
    for (B, C, H, W, dil, T) in [(2, 3, 40, 56, [1, 2, 4, 8, 12, 24], 10), (1, 5, 33, 47, [1, 2, 4, 8, 12], 3), (1, 2, 24, 24, [1, 3], 1)]:
        x = torch.from_numpy(mg.natural_image(rng, B, H, W))
        m = torch.from_numpy(mg.soft_mask(rng, B, C, H // 2, W // 2))   # exercises the align_corners resize too
        with torch.no_grad():
            want = wm.PAMR(T, dil)(x, m)
        assert torch.equal(tr.pamr(x, m, T, dil), want)
    for (H, W, n, thr, k) in [(64, 80, 9, 0.3, 41), (40, 40, 20, 0.1, 3), (33, 65, 4, 0.5, 7)]:
        heat, _ = mg.gaussian_heat(rng, H, W, n)
        heat = np.round(heat * 64) / 64
        want = mu.find_instance_center(torch.from_numpy(heat[None, None].copy()), thr, k, None)
        got = tr.find_instance_center(torch.from_numpy(heat[None, None].copy()), thr, k)
        assert torch.equal(got, want)
        off = torch.from_numpy((rng.standard_normal((1, 2, H, W)) * 20).astype(np.float32))
        if want.shape[0]:
            assert torch.equal(tr.group_pixels(want, off), mu.group_pixels(want, off))
    heat = torch.from_numpy(rng.random((2, 3, 48, 40)).astype(np.float32))
    ws, wy, wx = wu.peak_extract(heat, 5, 10)
    gs, gy, gx = tr.peak_extract(heat, 5, 10)
    assert np.array_equal(gs.numpy(), ws) and np.array_equal(gy.numpy(), wy) and np.array_equal(gx.numpy(), wx)
    assert torch.equal(tr.smoothing(heat, 3), wu.smoothing(heat, 3))
