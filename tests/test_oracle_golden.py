"""Pins the CPU oracle (oracle/cl4_oracle.c) to outputs of the unmodified reference.

Fixtures come from tests/golden/make_golden.py (reference run live in the build
container).  Tolerances: fp32 PAMR within rtol 1e-5 / atol 1e-7 of the reference
(the CUDA path is then held to the 1e-4 bar of BASELINE.md against the oracle);
centres, instance ids and positive peaks bit-exact.
"""
import numpy as np
import pytest

PAMR_CASES = ["pamr_d6", "pamr_d5", "pamr_tiny_d6", "pamr_1iter", "pamr_flat", "pamr_resize", "pamr_c21_64"]


@pytest.mark.parametrize("name", PAMR_CASES)
def test_pamr_matches_reference(golden, oracle, name):
    x, m, ref = golden[name + "__x"], golden[name + "__mask"], golden[name + "__out"]
    got = oracle.pamr(x, m, int(golden[name + "__T"]), golden[name + "__dil"].tolist())
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)


def test_pamr_weights_match_reference(golden, oracle):
    w = oracle.pamr_weights(golden["weights_d6__x"], [1, 2, 4, 8, 12, 24])
    np.testing.assert_allclose(w, golden["weights_d6__w"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(w.sum(1), 1.0, rtol=0, atol=1e-5)  # SURVEY §8c ③


def test_resize_identity_is_bit_exact(oracle):
    # SURVEY §8c ①
    m = np.random.default_rng(0).random((2, 3, 17, 23)).astype(np.float32)
    assert np.array_equal(oracle.resize_bilinear_ac(m, (17, 23)), m)


@pytest.mark.parametrize("name", ["peak_k15", "peak_k5", "peak_k3_neg", "peak_kat8"])
def test_peak_extract_matches_reference(golden, oracle, name):
    heat = golden[name + "__heat"]
    sc, ys, xs = oracle.peak_extract(heat, int(golden[name + "__kernel"]), int(golden[name + "__K"]))
    rs, ry, rx = golden[name + "__scores"], golden[name + "__ys"], golden[name + "__xs"]
    assert sc.dtype == np.float32 and ys.dtype == np.int32 and xs.dtype == np.int32  # SURVEY §8c ⑨
    assert np.array_equal(sc, rs)  # scores (sorted) are exact, fillers included
    # indices are defined where the score is positive and unique within its (b,c) row
    for b in range(sc.shape[0]):
        for c in range(sc.shape[1]):
            s = rs[b, c]
            uniq = np.array([(s == v).sum() == 1 for v in s]) & (s != 0)
            assert np.array_equal(ys[b, c][uniq], ry[b, c][uniq])
            assert np.array_equal(xs[b, c][uniq], rx[b, c][uniq])
            # tied scores: same coordinate set
            for v in np.unique(s[(s != 0) & ~uniq]):
                a = set(zip(ys[b, c][s == v].tolist(), xs[b, c][s == v].tolist()))
                r = set(zip(ry[b, c][s == v].tolist(), rx[b, c][s == v].tolist()))
                assert a == r
            # every reported location really holds the reported peak score or is a filler
            assert np.all((heat[b, c][ys[b, c], xs[b, c]] == sc[b, c]) | (sc[b, c] == 0))


def test_peak_kat8_values(golden, oracle):
    # SURVEY §8c ⑧
    sc, ys, xs = oracle.peak_extract(golden["peak_kat8__heat"], 5, 5)
    assert sc[0, 0].tolist() == [np.float32(0.9), np.float32(0.8), np.float32(0.8), 0.0, 0.0]
    assert list(zip(ys[0, 0][:3].tolist(), xs[0, 0][:3].tolist())) == [(5, 7), (20, 3), (20, 25)]


def test_find_instance_center_matches_reference(golden, oracle):
    n = int(golden["center__n"])
    assert n >= 10
    for i in range(n):
        thr, k, topk = golden[f"center_{i}__args"]
        topk = None if topk < 0 else int(topk)
        got = oracle.find_instance_center(golden[f"center_{i}__heat"], float(thr), int(k), topk)
        ref = golden[f"center_{i}__ctr"]
        assert got.dtype == np.int64 and got.shape == ref.shape, (i, got.shape, ref.shape)
        assert np.array_equal(got, ref), i


def test_find_instance_center_known_answers(oracle):
    # SURVEY §8c ⑥
    h = np.zeros((1, 1, 64, 64), np.float32)
    h[0, 0, 10, 10], h[0, 0, 10, 12], h[0, 0, 40, 40], h[0, 0, 41, 41], h[0, 0, 5, 60] = .9, .8, .5, .5, .05
    assert oracle.find_instance_center(h, 0.3, 3).tolist() == [[10, 10], [10, 12], [40, 40], [41, 41]]
    assert oracle.find_instance_center(h, 0.3, 5).tolist() == [[10, 10], [40, 40], [41, 41]]
    assert oracle.find_instance_center(h, 0.3, 41).tolist() == [[10, 10], [40, 40], [41, 41]]
    with pytest.raises(ValueError):
        oracle.find_instance_center(np.zeros((2, 1, 8, 8), np.float32))


def test_group_pixels_matches_reference(golden, oracle):
    n = int(golden["group__n"])
    for i in range(n):
        got = oracle.group_pixels(golden[f"group_{i}__ctr"], golden[f"group_{i}__off"])
        ref = golden[f"group_{i}__ids"]
        assert got.dtype == np.int64 and got.shape == ref.shape
        assert np.array_equal(got, ref), (i, int((got != ref).sum()))
    # SURVEY §8c ⑤
    ids = oracle.group_pixels(np.array([[1, 2], [1, 4]]), np.zeros((1, 2, 4, 8), np.float32))
    assert ids[0, 1].tolist() == [1, 1, 1, 1, 2, 2, 2, 2]


def test_get_instance_segmentation_matches_reference(golden, oracle):
    """All fixture cases, with and without centre clustering (beta > 0 uses OpenCV like the reference)."""
    n = int(golden["inst__n"])
    seen_beta = 0
    for i in range(n):
        thr, k, ignore, beta = golden[f"inst_{i}__args"]
        got, heat_after = oracle.get_instance_segmentation(golden[f"inst_{i}__fg"], golden[f"inst_{i}__heat"],
                                                           golden[f"inst_{i}__off"], float(thr), int(k), None,
                                                           bool(ignore), beta)
        assert np.array_equal(got, golden[f"inst_{i}__ids"]), i
        assert np.array_equal(heat_after, golden[f"inst_{i}__heat_after"]), i
        seen_beta += beta > 0
    assert n >= 5 and seen_beta >= 2


STENCIL_CASES = ["st_d6", "st_d5", "st_d1", "st_tiny", "st_odd"]


@pytest.mark.parametrize("name", STENCIL_CASES)
def test_helper_stencils_match_reference(golden_more, oracle, name):
    """LocalAffinity / Abs / Copy are single rounded operations: bit-exact.  LocalStDev: rtol 1e-5."""
    g = golden_more("stencils")
    x, dil = g[name + "__x"], g[name + "__dil"].tolist()
    for mode, key in ((0, "aff"), (1, "abs"), (2, "copy")):
        got = oracle.local_affinity(x, dil, mode)
        assert got.shape == g[f"{name}__{key}"].shape
        assert np.array_equal(got, g[f"{name}__{key}"]), (name, key)
    np.testing.assert_allclose(oracle.local_stdev(x, dil), g[name + "__std"], rtol=1e-5, atol=1e-7)


# --------------------------------------------------------------------------- callers
def _refine_args(g, ci):
    thr, kernel, beta, sigma, topk = g[f"refine_{ci}__args"]
    return dict(refine_thresh=float(thr), kernel=int(kernel), beta=float(beta), sigma=int(sigma)), (None if topk < 0 else int(topk))


def test_refine_label_generation_matches_reference(golden_more, oracle):
    """modules/utils.py:257-385: centre splats and offsets bit-exact, confidences to fp32 rounding
    (the reference's mean over the instance mask is a float reduction)."""
    g = golden_more("refine")
    for ci in range(int(g["n"])):
        k = f"refine_{ci}__"
        kw, topk = _refine_args(g, ci)
        r = oracle.labelgen.refine_label_generation(g[k + "seg"], g[k + "heat"], g[k + "off"], g[k + "label"], g[k + "gt"],
                                                    topk, **kw)
        assert np.array_equal(r["center"], g[k + "center"]), ci
        assert np.array_equal(r["offset"], g[k + "offset"]), ci
        assert np.array_equal(r["weight"] > 0, g[k + "weight"] > 0), ci
        np.testing.assert_allclose(r["weight"], g[k + "weight"], rtol=2e-6, atol=0)
    assert (g["refine_0__weight"] > 0).sum() > 1000 and (g["refine_3__weight"] > 0).sum() == 0


def test_refine_label_generation_with_point_matches_reference(golden_more, oracle):
    """modules/utils.py:388-460 (point supervision): offsets and weights bit-exact, including the reference's y != 0 and
    x != 0 point filter, invalid classes, duplicate points and the empty case."""
    g = golden_more("point")
    for ci in range(int(g["n"])):
        k = f"point_{ci}__"
        r = oracle.labelgen.refine_label_generation_with_point(g[k + "seg"], g[k + "points"], g[k + "off"], g[k + "label"], g[k + "gt"])
        assert np.array_equal(r["offset"], g[k + "offset"]), ci
        assert np.array_equal(r["weight"], g[k + "weight"]), ci
    assert (g["point_0__weight"] > 0).sum() > 1000 and (g["point_2__weight"] > 0).sum() == 0


def test_smoothing_and_pseudo_label_generation_match_reference(golden_more, oracle):
    g = golden_more("pseudo")
    for i in range(int(g["smooth__n"])):
        np.testing.assert_allclose(oracle.labelgen.smoothing(g[f"smooth_{i}__x"]), g[f"smooth_{i}__y"], rtol=1e-6, atol=1e-7)
    for ci in range(int(g["pseudo__n"])):
        k = f"pseudo_{ci}__"
        sigma = int(g[k + "sigma"])
        gg = oracle.labelgen.gaussian(sigma)
        assert np.array_equal(gg, g[k + "g"])
        pts = [[int(x), int(y), int(c), float(cf)] for x, y, c, cf in g[k + "points"]]
        c, o, w, n = oracle.labelgen.pseudo_label_generation(g[k + "gt"], pts, g[k + "label"], g[k + "center"].shape[0], sigma, gg)
        assert n == int(g[k + "match"])
        assert np.array_equal(c, g[k + "center"]) and np.array_equal(o, g[k + "offset"]) and np.array_equal(w, g[k + "weight"])


# --------------------------------------------------------------------------- phase-1 producers / consumers of PAMR
P1_CASES = ["p1_one", "p1_rect", "p1_same", "p1_voc"]


@pytest.mark.parametrize("name", P1_CASES)
def test_phase1_oracle_matches_reference(golden_more, name):
    """oracle/phase1.py against the reference's own denorm / interpolate / softmax / pseudo_gtmask outputs
    (tests/golden/make_golden_more.py, section phase1)."""
    from oracle import phase1 as p1
    g = golden_more("phase1")
    k = lambda s: g[f"{name}__{s}"]  # noqa: E731
    assert np.array_equal(p1.denorm(k("img")), k("denorm"))                      # two rounded operations: bit-exact
    np.testing.assert_allclose(p1.resize_bilinear_ac(k("denorm"), k("im").shape[-2:]), k("im"), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(p1.softmax_channels(k("logits")), k("soft"), rtol=2e-6, atol=1e-9)
    assert np.array_equal(p1.gate_labels(k("pamr"), k("l1h")), k("gated"))
    assert np.array_equal(p1.pseudo_gtmask(k("gated"), True, 0.6, 0.7, 0.2), k("pseudo"))
    assert np.array_equal(p1.pseudo_gtmask(k("soft"), False), k("pseudo_noamb"))


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_get_ins_map_oracle_golden(golden_more, oracle, ci):
    """oracle/labelgen.get_ins_map against the reference's get_ins_map (dataset/utils.py:795-902): plain, flip TTA +
    label cleaning + ignore, empty, clustered centres (score = seg_score^2), many centres per contour."""
    g = golden_more("insmap")
    k = f"insmap_{ci}__"
    thr, kern, beta, ign, flip, clean = g[k + "args"].tolist()
    seg_map, pl, pm, ps, off = oracle.labelgen.get_ins_map(g[k + "seg"], g[k + "center"], g[k + "offset"], g[k + "cls_label"],
                                                           tuple(int(v) for v in g[k + "target"]), thr, int(kern), beta,
                                                           bool(ign), bool(flip), bool(clean))
    assert np.array_equal(seg_map, g[k + "seg_map"])
    assert np.array_equal(pl, g[k + "pred_label"])
    shape = tuple(g[k + "pred_mask_shape"])
    want = np.unpackbits(g[k + "pred_mask"], axis=-1)[..., :shape[-1]].astype(bool)
    assert pm.shape == shape and np.array_equal(pm, want)
    np.testing.assert_allclose(ps, g[k + "pred_score"], rtol=2e-6, atol=0)
    assert np.array_equal(off, g[k + "offset_after"][0])


@pytest.mark.parametrize("name", ["cam_voc", "cam_odd", "cam_sized"])
def test_cam_chain_oracle_golden(golden_more, oracle, name):
    """oracle/labelgen.cam_normalize + upsample_bilinear against the reference's PeakGenerator.cam_normalize
    (wss/modules.py:425-434), smoothing, F.interpolate(align_corners=False) and peak_extract (train.py:426-436)."""
    g = golden_more("cam")
    k = name + "__"
    lg = oracle.labelgen
    norm = lg.cam_normalize(g[k + "cam"], tuple(g[k + "size"]), g[k + "label"])
    np.testing.assert_allclose(norm, g[k + "norm"], rtol=2e-6, atol=1e-7)
    sm = lg.smoothing(g[k + "norm"], 3)
    np.testing.assert_allclose(sm, g[k + "smooth"], rtol=2e-6, atol=1e-8)
    H, W = (int(v) for v in g[k + "image_size"])
    up = lg.upsample_bilinear(g[k + "smooth"], (H, W))
    if k + "up" in g.files:
        np.testing.assert_allclose(up, g[k + "up"], rtol=2e-6, atol=1e-7)
    kern, K = (int(v) for v in g[k + "kK"])
    from peaks_util import assert_peaks_equivalent
    assert_peaks_equivalent(oracle.peak_extract(up, kern, K), (g[k + "scores"], g[k + "ys"], g[k + "xs"]), up, kern)
