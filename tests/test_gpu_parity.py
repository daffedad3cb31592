"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in Python mirrors)
against the CPU oracle and the reference-generated golden fixtures.

Bars: refined masks allclose(rtol=1e-4, atol=1e-6) in fp32; centre lists,
peak indices and instance-id maps bit-exact.
"""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6
PAMR_CASES = ["pamr_d6", "pamr_d5", "pamr_tiny_d6", "pamr_1iter", "pamr_flat", "pamr_resize", "pamr_c21_64"]


@pytest.fixture(scope="module")
def cl4():
    import cl4wsis_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    cl4wsis_b200._lib.load()
    return cl4wsis_b200


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# --------------------------------------------------------------------------- helper stencils
@pytest.mark.parametrize("name", ["st_d6", "st_d5", "st_d1", "st_tiny", "st_odd"])
def test_helper_stencils_golden(cl4, golden_more, name):
    """wss/modules.py:17-119 drop-ins against reference outputs: differences / gathers bit-exact."""
    from cl4wsis_b200.wss import modules as wm
    g = golden_more("stencils")
    x, dil = cuda(g[name + "__x"]), g[name + "__dil"].tolist()
    for cls, key in ((wm.LocalAffinity, "aff"), (wm.LocalAffinityAbs, "abs"), (wm.LocalAffinityCopy, "copy")):
        got = cls(dil).cuda()(x).cpu().numpy()
        assert got.shape == g[f"{name}__{key}"].shape and got.dtype == np.float32
        assert np.array_equal(got, g[f"{name}__{key}"]), (name, key)
    got = wm.LocalStDev(dil).cuda()(x).cpu().numpy()
    assert got.shape == g[name + "__std"].shape
    np.testing.assert_allclose(got, g[name + "__std"], rtol=1e-5, atol=1e-7)


def test_helper_stencils_large_vs_oracle(cl4, oracle):
    from cl4wsis_b200.wss import modules as wm
    rng = np.random.default_rng(7)
    x = rng.random((2, 3, 200, 260)).astype(np.float32)
    dil = [1, 2, 4, 8, 12, 24]
    for mode, cls in enumerate((wm.LocalAffinity, wm.LocalAffinityAbs, wm.LocalAffinityCopy)):
        assert np.array_equal(cls(dil).cuda()(cuda(x)).cpu().numpy(), oracle.local_affinity(x, dil, mode))
    np.testing.assert_allclose(wm.LocalStDev(dil).cuda()(cuda(x)).cpu().numpy(), oracle.local_stdev(x, dil),
                               rtol=1e-5, atol=1e-7)
    # the reference's self-check (wss/modules.py:49-50): a modified stencil buffer trips the assert
    mod = wm.LocalAffinity(dil).cuda()
    mod.kernel[0, 0, 0, 0] = 5.0
    with pytest.raises(AssertionError):
        mod(cuda(x))


# --------------------------------------------------------------------------- PAMR
@pytest.mark.parametrize("name", PAMR_CASES)
def test_pamr_golden(cl4, golden, name):
    x, m, ref = golden[name + "__x"], golden[name + "__mask"], golden[name + "__out"]
    mod = cl4.PAMR(num_iter=int(golden[name + "__T"]), dilations=golden[name + "__dil"].tolist()).cuda()
    got = mod(cuda(x), cuda(m)).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B,C,H,W,dil,T", [
    (2, 21, 96, 80, [1, 2, 4, 8, 12, 24], 10),
    (1, 7, 33, 65, [1, 2, 4, 8, 12], 10),
    (2, 3, 50, 72, [1, 2, 4, 8, 12, 24], 10),    # partial tiles in both dimensions on the TMA path
    (1, 2, 32, 36, [1, 2, 4, 8, 12, 24], 3),     # one tile row: top and bottom frame from the same tile
    (1, 2, 200, 32, [1, 24], 5),
    (3, 2, 17, 130, [1, 2, 4, 8, 12, 24], 4),
    (1, 1, 5, 3, [1, 2, 4, 8, 12, 24], 10),      # every dilation exceeds the image
    (2, 3, 32, 32, [1, 2, 4, 8, 12], 10),        # the trainer's feature-resolution regime
    (1, 81, 56, 56, [1, 2, 4, 8, 12], 10),       # coco-voc feature resolution, 81 classes
    (1, 4, 40, 40, [3], 2),
    (1, 4, 40, 40, [1, 2, 3, 4, 5, 6, 7, 8], 2),
    (1, 3, 24, 24, [1, 2, 4], 0),                # num_iter = 0 returns the (resized) mask
])
def test_pamr_oracle_random(cl4, oracle, B, C, H, W, dil, T):
    rng = np.random.default_rng(B * 1000 + C * 100 + H + W)
    x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    want = oracle.pamr(x, m, T, dil)
    got = cl4.PAMR(T, dil).cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("path", ["fused", "tma", "v1"])
@pytest.mark.parametrize("B,C,H,W,dil,T", [
    (16, 21, 32, 32, [1, 2, 4, 8, 12], 10),      # VOC phase-1 call: train.py:81, :379 at crop 512 / stride 16
    (2, 81, 56, 56, [1, 2, 4, 8, 12], 10),       # coco-voc feature resolution: 4 tiles, weight reload on-chip
    (3, 5, 64, 64, [1, 2, 4, 8, 12, 24], 7),     # largest fused map, odd iteration count
    (2, 4, 33, 20, [1, 2, 4, 8, 12, 24], 2),     # two tile rows, one column, W % 4 == 0
    (1, 3, 40, 36, [2, 5, 24], 3),               # runtime dilation offsets
    (1, 2, 32, 64, [1, 2, 4, 8, 12, 24], 1),     # single iteration: straight to the output
    (2, 5, 57, 64, [1, 2, 4, 8, 12], 3),         # frame of 12 (no dilation above 12): partial second tile row, odd class count
    (1, 7, 64, 33, [12], 4),                     # frame of 12, runtime dilation offsets, one pixel column in the second tile
    (1, 2, 33, 33, [5, 12], 3),                  # frame of 12, 2 x 2 tiles of which three hold one row / column
    (1, 3, 48, 40, [13], 2),                     # one above the small frame: back to the frame of 24
])
def test_pamr_small_map_paths_agree(cl4, oracle, monkeypatch, path, B, C, H, W, dil, T):
    """The three sweep implementations (all iterations on-chip / TMA-staged / register-L1) on the
    small maps the trainer really feeds PAMR, each against the oracle."""
    if path != "fused":
        monkeypatch.setenv("CL4_SWEEP", path)
    rng = np.random.default_rng(B + 10 * C + 100 * H + W)
    x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    got = cl4.PAMR(T, dil).cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, oracle.pamr(x, m, T, dil), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("mode", ["lattice", "lattice1", "nolattice"])
@pytest.mark.parametrize("B,C,H,W,T", [
    (2, 21, 96, 80, 10),     # partial tiles in x, 3 x 3 tiles
    (2, 3, 50, 72, 10),      # partial tiles in both dimensions
    (1, 2, 200, 36, 5),      # tall and narrow: many tile rows per CTA
    (3, 2, 68, 132, 4),      # more tiles than one wave of classes; staggered class phase
    (1, 1, 72, 68, 1),       # single class, single iteration: straight to the output
    (5, 30, 160, 160, 2),    # 125 tiles x 30 classes: every CTA switches tiles (weight reload on the fly)
    (2, 4, 36, 40, 3),       # even class count, 2 x 2 partial tiles
])
@pytest.mark.parametrize("dil", [[1, 2, 4, 8, 12, 24], [1, 2, 4, 8, 12]])
def test_pamr_lattice_sweep(cl4, oracle, monkeypatch, mode, dil, B, C, H, W, T):
    """The class-pair lattice sweep (pamr_duo.cu, the default: "lattice"), the one-class lattice sweep (pamr_lattice.cu,
    "lattice1") and the 4-pixel TMA sweep on the class-default dilation set (wss/modules.py:125) and on the trainer's
    (train.py:81), each against the oracle.  Odd and even class counts: an odd count runs one dummy class."""
    monkeypatch.setenv("CL4_SWEEP", mode)
    rng = np.random.default_rng(B * 1000 + C * 100 + H + W)
    x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    got = cl4.PAMR(T, dil).cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, oracle.pamr(x, m, T, dil), rtol=RTOL, atol=ATOL)


def test_pamr_fused_odd_width_and_resize(cl4, oracle):
    """The fused path has no W % 4 requirement and composes with the bilinear resize of :134."""
    rng = np.random.default_rng(11)
    x = rng.random((2, 3, 29, 37)).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((2, 6, 8, 10)).astype(np.float32)).softmax(1).numpy()
    got = cl4.PAMR(10, [1, 2, 4, 8, 12]).cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, oracle.pamr(x, m, 10, [1, 2, 4, 8, 12]), rtol=RTOL, atol=ATOL)


def test_pamr_weights_and_single_sweep(cl4, oracle, golden):
    """The two kernels separately, through the C ABI."""
    lib, L = cl4._lib.load(), cl4._lib
    x = golden["weights_d6__x"]
    dil = [1, 2, 4, 8, 12, 24]
    B, K, H, W = x.shape
    xd = cuda(x)
    w = torch.empty((B, 48, H, W), dtype=torch.float32, device="cuda")
    L.check(lib.cl4_pamr_weights(L.ptr(xd), L.ptr(w), B, K, H, W, L.int_array(dil), 6, L.stream_ptr()), "weights")
    np.testing.assert_allclose(w.cpu().numpy(), golden["weights_d6__w"], rtol=1e-4, atol=1e-7)
    m = np.random.default_rng(3).random((B, 5, H, W)).astype(np.float32)
    md, out = cuda(m), torch.empty((B, 5, H, W), dtype=torch.float32, device="cuda")
    L.check(lib.cl4_pamr_sweep(L.ptr(w), L.ptr(md), L.ptr(out), B, 5, H, W, L.int_array(dil), 6, L.stream_ptr()), "sweep")
    # one sweep == PAMR with num_iter=1
    np.testing.assert_allclose(out.cpu().numpy(), oracle.pamr(x, m, 1, dil), rtol=RTOL, atol=ATOL)


def test_pamr_module_state_and_errors(cl4):
    mod = cl4.PAMR(num_iter=10, dilations=[1, 2, 4, 8, 12])
    sd = mod.state_dict()
    assert sorted(sd) == ["aff_m.kernel", "aff_std.kernel", "aff_x.kernel"]
    assert tuple(sd["aff_x.kernel"].shape) == (8, 1, 3, 3) and tuple(sd["aff_std.kernel"].shape) == (9, 1, 3, 3)
    assert len(list(mod.parameters())) == 0
    with pytest.raises(RuntimeError):
        mod(torch.rand(1, 3, 8, 8), torch.rand(1, 2, 8, 8))  # CPU tensors: no fallback
    with pytest.raises(TypeError):
        mod.cuda()(torch.rand(1, 3, 8, 8, device="cuda").half(), torch.rand(1, 2, 8, 8, device="cuda"))
    with pytest.raises(NotImplementedError):
        cl4.PAMR(1, list(range(1, 10))).cuda()(torch.rand(1, 3, 8, 8, device="cuda"), torch.rand(1, 2, 8, 8, device="cuda"))


def test_pamr_full_size_properties(cl4):
    """BASELINE config 2 shape (B cut to 2): channel sums stay 1 (SURVEY §8c ③), linearity in
    the mask, and a constant mask is a fixed point (weights sum to 1)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    B, C, H, W = 2, 21, 512, 512
    x = (torch.randint(0, 256, (B, 3, H, W), generator=g).float() / 255).cuda()
    m1 = torch.randn((B, C, H, W), generator=g).softmax(1).cuda()
    m2 = torch.rand((B, C, H, W), generator=g).cuda()
    mod = cl4.PAMR(10, [1, 2, 4, 8, 12, 24]).cuda()
    o1, o2 = mod(x, m1), mod(x, m2)
    assert torch.allclose(o1.sum(1), torch.ones_like(o1[:, 0]), rtol=0, atol=2e-5)
    o12 = mod(x, 0.25 * m1 + 0.75 * m2)
    assert torch.allclose(o12, 0.25 * o1 + 0.75 * o2, rtol=1e-4, atol=1e-6)
    const = torch.full((B, C, H, W), 0.37, device="cuda")
    assert torch.allclose(mod(x, const), const, rtol=0, atol=1e-5)
    assert float(o1.min()) >= 0.0 and float(o1.max()) <= 1.0 + 1e-5


def test_pamr_full_size_tile_vs_oracle(cl4, oracle):
    """One 512x512 image, 3 classes, against the oracle (finishes in seconds on CPU)."""
    rng = np.random.default_rng(77)
    x = (rng.integers(0, 256, (1, 3, 512, 512)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((1, 3, 512, 512)).astype(np.float32)).softmax(1).numpy()
    want = oracle.pamr(x, m, 10, [1, 2, 4, 8, 12, 24])
    got = cl4.PAMR().cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B,C,H,W,dil", [
    (1, 21, 512, 512, [1, 2, 4, 8, 12, 24]),   # the headline shape's class count on a full-size map (config 2)
    (1, 81, 512, 512, [1, 2, 4, 8, 12, 24]),   # config 3 (COCO-to-VOC): 41 class pairs per tile
    (1, 81, 512, 512, [1, 2, 4, 8, 12]),       # the trainer's dilation set at the same shape
    (1, 3, 1024, 1024, [1, 2, 4, 8, 12, 24]),  # config 4 resolution: 1024 tiles, 7 tiles per CTA
])
def test_pamr_large_shapes_vs_oracle(cl4, oracle, B, C, H, W, dil):
    """The class-pair sweep where it is benchmarked: the staggered pair phase (s0 = blockIdx * Cp / grid) and the
    weight refill at tile switches depend on C and on the number of tiles per CTA (wss/modules.py:147-149)."""
    rng = np.random.default_rng(C * 7 + H)
    x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    want = oracle.pamr(x, m, 10, dil)
    got = cl4.PAMR(10, dil).cuda()(cuda(x), cuda(m)).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


# --------------------------------------------------------------------------- peak_extract
def _check_peaks(got, want, heat):
    sc, ys, xs = got
    ws, wy, wx = want
    assert sc.dtype == np.float32 and ys.dtype == np.int32 and xs.dtype == np.int32
    assert np.array_equal(sc, ws)
    # same tie rule as the oracle (score desc, flat index asc): indices are exact everywhere
    assert np.array_equal(ys, wy) and np.array_equal(xs, wx)


@pytest.mark.parametrize("name", ["peak_k15", "peak_k5", "peak_k3_neg", "peak_kat8"])
def test_peak_extract_golden(cl4, oracle, golden, name):
    heat = golden[name + "__heat"]
    k, K = int(golden[name + "__kernel"]), int(golden[name + "__K"])
    got = cl4.peak_extract(cuda(heat), kernel=k, K=K)
    _check_peaks(got, oracle.peak_extract(heat, k, K), heat)
    # against the reference itself: scores exact; indices where the score is positive and unique
    rs, ry, rx = golden[name + "__scores"], golden[name + "__ys"], golden[name + "__xs"]
    assert np.array_equal(got[0], rs)
    for b in range(rs.shape[0]):
        for c in range(rs.shape[1]):
            s = rs[b, c]
            uniq = np.array([(s == v).sum() == 1 for v in s]) & (s != 0)
            assert np.array_equal(got[1][b, c][uniq], ry[b, c][uniq])
            assert np.array_equal(got[2][b, c][uniq], rx[b, c][uniq])


@pytest.mark.parametrize("B,C,H,W,kernel,K", [(2, 20, 128, 96, 15, 25), (1, 3, 70, 33, 5, 40), (1, 2, 64, 64, 41, 100),
                                              (1, 1, 9, 9, 3, 81), (1, 2, 200, 300, 15, 256)])
def test_peak_extract_oracle_random(cl4, oracle, B, C, H, W, kernel, K):
    rng = np.random.default_rng(H * W + K)
    heat = rng.random((B, C, H, W)).astype(np.float32)
    heat[:, :, : H // 2] = np.round(heat[:, :, : H // 2] * 16) / 16  # exact ties and plateaus
    heat[:, 0, H // 2:, : W // 2] = 0.0                              # zero fillers
    _check_peaks(cl4.peak_extract(cuda(heat), kernel=kernel, K=K), oracle.peak_extract(heat, kernel, K), heat)


def test_peak_extract_errors(cl4):
    h = torch.rand(1, 1, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        cl4.peak_extract(h, kernel=4, K=3)
    with pytest.raises(RuntimeError):
        cl4.peak_extract(h, kernel=3, K=65)
    with pytest.raises(RuntimeError):
        cl4.peak_extract(torch.rand(1, 1, 8, 8), kernel=3, K=3)


@pytest.mark.parametrize("B,C,H,W,kernel,K", [(1, 2, 64, 96, 5, 300), (2, 1, 40, 40, 3, 1000), (1, 1, 33, 31, 15, 1023),
                                              (1, 3, 128, 128, 41, 257)])
def test_peak_extract_large_K(cl4, oracle, B, C, H, W, kernel, K):
    """The reference accepts any K <= H*W (torch.topk, wss/utils.py:15); K > 256 is selected in rounds of 256."""
    rng = np.random.default_rng(K)
    heat = rng.random((B, C, H, W)).astype(np.float32)
    heat[0, 0, :8] = np.round(heat[0, 0, :8] * 4) / 4        # plateaus: ties on the score
    heat[0, 0, 20:24] = -heat[0, 0, 20:24]                    # negative peaks sort below the zeros
    _check_peaks(cl4.peak_extract(cuda(heat), kernel=kernel, K=K), oracle.peak_extract(heat, kernel, K), heat)


@pytest.mark.parametrize("name", ["cam_voc", "cam_odd", "cam_sized"])
def test_cam_chain_golden(cl4, golden_more, oracle, name):
    """train.py:426-436 (cam_normalize -> smoothing -> F.interpolate -> peak_extract) against reference outputs.  With
    size = None (the trainer's call) cam_normalize is bit-exact; a real resize and the fused up-sampling inside the peak
    loader agree to fp32 rounding (ATen's own CPU and CUDA kernels differ by as much), so the peak lists are compared
    with tests/peaks_util.assert_peaks_equivalent: same pixels and scores except pixels that tie with their window
    maximum to the last bits."""
    from cl4wsis_b200.wss import utils as wu
    from peaks_util import assert_peaks_equivalent
    g = golden_more("cam")
    k = name + "__"
    cam, label = cuda(g[k + "cam"]), cuda(g[k + "label"])
    size = tuple(int(v) for v in g[k + "size"])
    norm = wu.cam_normalize(cam, size, label)
    if size == tuple(cam.shape[-2:]):
        assert np.array_equal(norm.cpu().numpy(), g[k + "norm"])
        assert np.array_equal(wu.cam_normalize(cam, None, label).cpu().numpy(), g[k + "norm"])
    else:
        np.testing.assert_allclose(norm.cpu().numpy(), g[k + "norm"], rtol=2e-6, atol=1e-7)
    kern, K = (int(v) for v in g[k + "kK"])
    img = tuple(int(v) for v in g[k + "image_size"])
    sc, ys, xs = (t.cpu().numpy() for t in wu.peak_extract_device(wu.smoothing(cuda(g[k + "norm"]), 3), kern, K, upsample_to=img))
    up = oracle.labelgen.upsample_bilinear(g[k + "smooth"], img)
    assert_peaks_equivalent((sc, ys, xs), (g[k + "scores"], g[k + "ys"], g[k + "xs"]), up, kern)
    strong = g[k + "scores"] >= 0.25          # the peaks the trainer keeps have conf >= pseudo_thresh (train.py:456)
    np.testing.assert_allclose(sc[strong], g[k + "scores"][strong], rtol=1e-5)
    assert np.array_equal(ys[strong], g[k + "ys"][strong]) and np.array_equal(xs[strong], g[k + "xs"][strong])
    if size == tuple(cam.shape[-2:]):  # the whole chain from the raw CAM in one call
        sc2, ys2, xs2 = (t.cpu().numpy() for t in wu.cam_peaks(cam, label, img, 3, kern, K))
        assert np.array_equal(sc2, sc) and np.array_equal(ys2, ys) and np.array_equal(xs2, xs)


def test_peak_extract_upsampled_vs_materialised(cl4, fp32_convs):
    """The trainer's shapes (B16, 20 classes, 32x32 -> 512x512, kernel 15, K 25): peaks of the on-the-fly up-sampling
    against peak_extract of the map F.interpolate writes on this GPU."""
    from cl4wsis_b200.wss import utils as wu
    g = torch.Generator(device="cpu").manual_seed(31)
    small = torch.rand((16, 20, 32, 32), generator=g).cuda()
    small = wu.smoothing(small * (small > 0.5), 3)
    up = torch.nn.functional.interpolate(small, size=(512, 512), mode="bilinear", align_corners=False)
    ws, wy, wx = (t.cpu().numpy() for t in wu.peak_extract_device(up, 15, 25))
    s, y, x = (t.cpu().numpy() for t in wu.peak_extract_device(small, 15, 25, upsample_to=(512, 512)))
    from peaks_util import assert_peaks_equivalent
    assert_peaks_equivalent((s, y, x), (ws, wy, wx), up.cpu().numpy(), 15)


# --------------------------------------------------------------------------- find_instance_center
def test_find_instance_center_golden(cl4, golden):
    for i in range(int(golden["center__n"])):
        thr, k, topk = golden[f"center_{i}__args"]
        topk = None if topk < 0 else int(topk)
        heat = cuda(golden[f"center_{i}__heat"])
        before = heat.clone()
        with redirect_stdout(io.StringIO()) as so:
            got = cl4.find_instance_center(heat, float(thr), int(k), topk)
        ref = golden[f"center_{i}__ctr"]
        assert got.dtype == torch.int64 and got.is_cuda and tuple(got.shape) == ref.shape, i
        assert np.array_equal(got.cpu().numpy(), ref), i
        assert torch.equal(heat, before)  # input not modified
        if f"center_{i}__printed" in golden.files:
            assert int(so.getvalue().strip()) == int(golden[f"center_{i}__printed"])  # SURVEY §8c ⑦


@pytest.mark.parametrize("H,W,k,thr", [(512, 512, 41, 0.3), (100, 333, 3, 0.1), (65, 31, 7, 0.5), (33, 2, 5, 0.2),
                                       (2, 200, 9, 0.2), (256, 256, 1, 0.6)])
def test_find_instance_center_oracle_random(cl4, oracle, H, W, k, thr):
    rng = np.random.default_rng(H + 7 * W + k)
    heat = rng.random((1, 1, H, W)).astype(np.float32)
    heat = np.round(heat * 32) / 32  # ties
    heat[0, 0, : H // 3] = np.nan if H > 64 else heat[0, 0, : H // 3]  # NaN rows suppress their windows
    want = oracle.find_instance_center(heat, thr, k)
    got = cl4.find_instance_center(cuda(heat), thr, k).cpu().numpy()
    assert np.array_equal(got, want)


def test_find_instance_center_many_centres(cl4, oracle):
    # more centres than the wrapper's first capacity guess (4096)
    heat = np.zeros((1, 1, 256, 256), np.float32)
    heat[0, 0, ::2, ::2] = 0.5
    got = cl4.find_instance_center(cuda(heat), 0.3, 1).cpu().numpy()
    assert got.shape == (128 * 128, 2) and np.array_equal(got, oracle.find_instance_center(heat, 0.3, 1))


def test_find_instance_center_errors(cl4):
    with pytest.raises(ValueError, match="batch size = 1"):
        cl4.find_instance_center(torch.zeros(2, 1, 8, 8, device="cuda"))
    with pytest.raises(AssertionError):
        cl4.find_instance_center(torch.zeros(1, 2, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        cl4.find_instance_center(torch.zeros(1, 1, 8, 8, device="cuda"), 0.1, 4)


# --------------------------------------------------------------------------- group_pixels
def test_group_pixels_golden(cl4, golden):
    for i in range(int(golden["group__n"])):
        got = cl4.group_pixels(cuda(golden[f"group_{i}__ctr"]), cuda(golden[f"group_{i}__off"]))
        ref = golden[f"group_{i}__ids"]
        assert got.dtype == torch.int64 and tuple(got.shape) == ref.shape
        assert np.array_equal(got.cpu().numpy(), ref), i


@pytest.mark.parametrize("H,W,Kc", [(512, 512, 5), (1024, 1024, 200), (33, 31, 7), (17, 19, 1), (64, 50, 3000), (5, 3, 2)])
def test_group_pixels_oracle_random(cl4, oracle, H, W, Kc):
    rng = np.random.default_rng(H * 3 + W + Kc)
    ctr = np.stack([rng.integers(0, H, Kc), rng.integers(0, W, Kc)], 1).astype(np.int64)
    off = (rng.standard_normal((1, 2, H, W)) * 15).astype(np.float32)
    off[0, :, : H // 4] = np.round(off[0, :, : H // 4])  # integer offsets: exact distance ties
    want = oracle.group_pixels(ctr, off)
    got = cl4.group_pixels(cuda(ctr), cuda(off)).cpu().numpy()
    assert np.array_equal(got, want), int((got != want).sum())
    assert got.min() >= 1 and got.max() <= Kc


def test_group_pixels_errors(cl4):
    with pytest.raises(ValueError, match="batch size = 1"):
        cl4.group_pixels(torch.zeros(1, 2, dtype=torch.long, device="cuda"), torch.zeros(2, 2, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        cl4.group_pixels(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, 8, 8))


# --------------------------------------------------------------------------- get_instance_segmentation
def test_get_instance_segmentation_golden(cl4, golden):
    """Every fixture case, with and without centre clustering; the in-place marking of merged cluster
    centres in ctr_hmp (modules/utils.py:583,591) is part of the contract."""
    n, seen_beta = int(golden["inst__n"]), 0
    for i in range(n):
        thr, k, ignore, beta = golden[f"inst_{i}__args"]
        hm = cuda(golden[f"inst_{i}__heat"])
        got = cl4.get_instance_segmentation(cuda(golden[f"inst_{i}__fg"]), hm, cuda(golden[f"inst_{i}__off"]),
                                            threshold=float(thr), nms_kernel=int(k), top_k=None, ignore=bool(ignore),
                                            beta=beta)
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), golden[f"inst_{i}__ids"]), i
        assert np.array_equal(hm.cpu().numpy(), golden[f"inst_{i}__heat_after"]), i
        seen_beta += beta > 0
    assert n >= 9 and seen_beta >= 6


@pytest.mark.parametrize("H,W,p,beta", [(64, 80, 0.5, 5), (200, 333, 0.3, 3.0), (37, 41, 0.7, 20), (512, 512, 0.45, 5),
                                        (16, 16, 0.97, 21)])
def test_cluster_peaks_matches_opencv(cl4, oracle, H, W, p, beta):
    """GPU connected components vs the reference's cv2 path: same components, same order, same
    truncated centroids — including OpenCV's label 0 (the complement) when its area passes the filter."""
    from cl4wsis_b200.cluster import cluster_peaks
    rng = np.random.default_rng(H * W)
    off = (rng.standard_normal((1, 2, H, W)) * (2.5 / np.sqrt(-2 * np.log(1 - p + 1e-9)))).astype(np.float32)
    fg = rng.random((1, H, W)) > 0.1
    want = oracle.cluster_peaks(off[0], fg[0], beta=beta)
    got = cluster_peaks(cuda(off), cuda(fg), beta=beta)
    assert got.dtype == np.int32 and got.shape == want.reshape(-1, 2).shape
    assert np.array_equal(got, want.reshape(-1, 2))


# --------------------------------------------------------------------------- batched step
def test_pseudo_label_step_matches_per_image_oracle(cl4, oracle):
    rng = np.random.default_rng(2024)
    B, C, H, W = 3, 4, 96, 128
    x = (rng.integers(0, 256, (B, 3, H, W)) / 255.0).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    heat = np.zeros((B, 1, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for b in range(B):
        for _ in range(b * 3):  # image 0 has no centre at all
            cy, cx, a = rng.integers(0, H), rng.integers(0, W), rng.uniform(0.4, 1.0)
            heat[b, 0] = np.maximum(heat[b, 0], a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72.0))
    off = (rng.standard_normal((B, 2, H, W)) * 10).astype(np.float32)
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=10, dilations=[1, 2, 4, 8, 12, 24], threshold=0.3, nms_kernel=41,
                               max_centers=64)
    refined, ids, counts, centers = step.run(cuda(x), cuda(m), cuda(heat), cuda(off))
    torch.cuda.synchronize()
    np.testing.assert_allclose(refined.cpu().numpy(), oracle.pamr(x, m, 10, [1, 2, 4, 8, 12, 24]), rtol=RTOL, atol=ATOL)
    for b in range(B):
        ctr = oracle.find_instance_center(heat[b:b + 1], 0.3, 41)
        assert int(counts[b]) == ctr.shape[0]
        assert np.array_equal(centers[b, : ctr.shape[0]].cpu().numpy(), ctr)
        if ctr.shape[0] == 0:
            assert int(ids[b].abs().sum()) == 0  # ignore=True: zeros (modules/utils.py:597-598)
        else:
            assert np.array_equal(ids[b].cpu().numpy(), oracle.group_pixels(ctr, off[b:b + 1])[0])
    assert not bool(step.overflowed().any())
    step.assert_no_overflow()


def test_pseudo_label_step_reports_truncated_centre_lists(cl4, oracle):
    """More centres than max_centers: counts keeps the true total, the stored list is the first max_centers in nonzero
    order, and the overflow is reported instead of passing silently."""
    rng = np.random.default_rng(7)
    B, C, H, W = 2, 2, 64, 64
    x = rng.random((B, 3, H, W)).astype(np.float32)
    m = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).softmax(1).numpy()
    heat = np.zeros((B, 1, H, W), np.float32)
    heat[1, 0, 4::8, 4::8] = rng.uniform(0.5, 1.0, (8, 8)).astype(np.float32)  # 64 isolated maxima at k = 3
    heat[0, 0, 10, 10] = 0.9
    off = np.zeros((B, 2, H, W), np.float32)
    step = cl4.PseudoLabelStep(B, C, H, W, num_iter=1, threshold=0.3, nms_kernel=3, max_centers=16)
    _, _, counts, centers = step.run(cuda(x), cuda(m), cuda(heat), cuda(off))
    assert counts.tolist() == [1, 64]
    assert step.overflowed().tolist() == [False, True]
    ctr = oracle.find_instance_center(heat[1:2], 0.3, 3)
    assert np.array_equal(centers[1].cpu().numpy(), ctr[:16])
    with pytest.raises(OverflowError):
        step.assert_no_overflow()


# --------------------------------------------------------------------------- refine_label_generation
class _Args:
    def __init__(self, refine_thresh, kernel, beta, sigma):
        self.refine_thresh, self.kernel, self.beta, self.sigma = refine_thresh, kernel, beta, sigma


def _refine_case(g, ci):
    k = f"refine_{ci}__"
    thr, kernel, beta, sigma, topk = g[k + "args"]
    args = _Args(float(thr), int(kernel), float(beta), int(sigma))
    ins = [cuda(g[k + n]) for n in ("seg", "heat", "off", "label", "gt")]
    return k, ins, (None if topk < 0 else int(topk)), args


def _check_refine(r, g, k):
    assert np.array_equal(r["center"].cpu().numpy(), g[k + "center"])
    assert np.array_equal(r["offset"].cpu().numpy(), g[k + "offset"])
    w = r["weight"].cpu().numpy()
    assert np.array_equal(w > 0, g[k + "weight"] > 0)
    np.testing.assert_allclose(w, g[k + "weight"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_refine_label_generation_golden(cl4, golden_more, ci):
    """Batched device path against the reference's outputs: centre splats and offsets bit-exact,
    confidences to fp32 rounding."""
    from cl4wsis_b200.modules import utils as mu
    g = golden_more("refine")
    k, ins, topk, args = _refine_case(g, ci)
    r, status = mu.refine_label_generation_device(*ins, topk, args)
    if ci == 2:  # noisy heat + 5x5 NMS: > 64 centres in one contour -> status bit 1 -> per-contour path
        assert int(status.item()) == 2
    else:
        assert int(status.item()) == 0
        _check_refine(r, g, k)
    _check_refine(mu.refine_label_generation(*ins, topk, args), g, k)


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_refine_label_generation_per_contour_golden(cl4, golden_more, ci):
    """The exact per-contour path (used on overflow) against the same fixtures."""
    from cl4wsis_b200.modules import utils as mu
    g = golden_more("refine")
    k, ins, topk, args = _refine_case(g, ci)
    _check_refine(mu.refine_label_generation_per_contour(*ins, topk, args), g, k)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_refine_label_generation_with_point_golden(cl4, golden_more, ci):
    """modules/utils.py:388-460 against the reference fixtures: bit-exact offsets and weights."""
    from cl4wsis_b200.modules import utils as mu
    g = golden_more("point")
    k = f"point_{ci}__"
    r = mu.refine_label_generation_with_point(cuda(g[k + "seg"]), cuda(g[k + "points"]), cuda(g[k + "off"]), cuda(g[k + "label"]),
                                              cuda(g[k + "gt"]), None)
    assert np.array_equal(r["offset"].cpu().numpy(), g[k + "offset"])
    assert np.array_equal(r["weight"].cpu().numpy(), g[k + "weight"])


def test_refine_label_generation_with_point_random_vs_oracle(cl4, oracle):
    from cl4wsis_b200.modules import utils as mu
    rng = np.random.default_rng(11)
    B, C, M, H, W = 3, 20, 12, 200, 333
    gt = rng.integers(0, C + 2, (B, H // 8 + 1, W // 8 + 1)).repeat(8, 1).repeat(8, 2)[:, :H, :W].astype(np.int64)  # incl. label C + 1
    lab = (rng.random((B, C)) < 0.7).astype(np.float32)
    pts = np.zeros((B, C, M, 2), np.int64)
    for b in range(B):
        for c in range(C):
            n = rng.integers(0, M + 1)
            pts[b, c, :n, 0] = rng.integers(0, H, n)     # a zero coordinate now and then: filtered
            pts[b, c, :n, 1] = rng.integers(0, W, n)
    off = (rng.standard_normal((B, 2, H, W)) * 20).astype(np.float32)
    seg = np.zeros((B, C + 1, H, W), np.float32)
    r = mu.refine_label_generation_with_point(cuda(seg), cuda(pts), cuda(off), cuda(lab), cuda(gt), None)
    o = oracle.labelgen.refine_label_generation_with_point(seg, pts, off, lab, gt)
    assert np.array_equal(r["offset"].cpu().numpy(), o["offset"])
    assert np.array_equal(r["weight"].cpu().numpy(), o["weight"])
    assert 0 < (o["weight"] > 0).mean() < 1


def test_refine_label_generation_overflow_falls_back(cl4, golden_more):
    """A degenerate top_k (>= centres of a contour) trips the status word; the drop-in then answers
    through the per-contour path, which reproduces the reference's degenerate branch."""
    from cl4wsis_b200.modules import utils as mu
    g = golden_more("refine")
    k, ins, _, args = _refine_case(g, 0)
    _, status = mu.refine_label_generation_device(*ins, 1, args)
    assert int(status.item()) & 8
    with redirect_stdout(io.StringIO()):
        a = mu.refine_label_generation(*ins, 1, args)
        b = mu.refine_label_generation_per_contour(*ins, 1, args)
    for key in ("center", "offset", "weight"):
        assert torch.equal(a[key], b[key])


def test_refine_label_generation_random_vs_oracle(cl4, oracle):
    """Larger random scenes (512x512, 20 classes): device path against the numpy/OpenCV oracle."""
    from cl4wsis_b200.modules import utils as mu
    rng = np.random.default_rng(77)
    B, C, H, W = 2, 20, 512, 512
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    seg = rng.standard_normal((B, C + 1, H, W)).astype(np.float32)
    heat = (0.05 * rng.random((B, C, H, W))).astype(np.float32)
    off = (0.3 * rng.standard_normal((B, 2, H, W)) + 40).astype(np.float32)
    gt = np.zeros((B, H, W), np.int64)
    lab = np.zeros((B, C), np.float32)
    for b in range(B):
        for _ in range(14):
            cls = int(rng.integers(0, C)); cy, cx = int(rng.integers(20, H - 20)), int(rng.integers(20, W - 20))
            ry, rx = int(rng.integers(4, 60)), int(rng.integers(4, 60))
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
            gt[b][m] = cls + 1
            if rng.random() < 0.85:
                lab[b, cls] = 1
            if rng.random() < 0.8:
                heat[b, cls] = np.maximum(heat[b, cls], rng.uniform(0.25, 0.95) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72).astype(np.float32))
            off[b, 0][m] = (cy - yy)[m] + 0.3 * rng.standard_normal(int(m.sum()))
            off[b, 1][m] = (cx - xx)[m] + 0.3 * rng.standard_normal(int(m.sum()))
        for c in range(C + 1):
            seg[b, c][gt[b] == c] += 3
    args = _Args(0.3, 41, 3.0, 6)
    want = oracle.labelgen.refine_label_generation(seg, heat, off, lab, gt, 10000, refine_thresh=0.3, kernel=41, beta=3.0, sigma=6)
    r, status = mu.refine_label_generation_device(cuda(seg), cuda(heat), cuda(off), cuda(lab), cuda(gt), 10000, args)
    assert int(status.item()) == 0
    assert (want["weight"] > 0).sum() > 10000
    assert np.array_equal(r["center"].cpu().numpy(), want["center"])
    assert np.array_equal(r["offset"].cpu().numpy(), want["offset"])
    np.testing.assert_allclose(r["weight"].cpu().numpy(), want["weight"], rtol=2e-6, atol=0)


# --------------------------------------------------------------------------- smoothing / pseudo labels
def test_smoothing_golden(cl4, golden_more, oracle):
    from cl4wsis_b200.wss.utils import smoothing
    g = golden_more("pseudo")
    for i in range(int(g["smooth__n"])):
        np.testing.assert_allclose(smoothing(cuda(g[f"smooth_{i}__x"])).cpu().numpy(), g[f"smooth_{i}__y"], rtol=1e-6, atol=1e-7)
    x = np.random.default_rng(1).random((2, 20, 64, 64)).astype(np.float32)
    for k in (3, 5):
        np.testing.assert_allclose(smoothing(cuda(x), k).cpu().numpy(), oracle.labelgen.smoothing(x, k), rtol=1e-6, atol=1e-7)


def _peaks_from_points(points, C, K):
    conf = np.zeros((1, C, K), np.float32); ys = np.zeros((1, C, K), np.int32); xs = np.zeros((1, C, K), np.int32)
    for c in range(C):
        pts = sorted([p for p in points if int(p[2]) == c], key=lambda p: -p[3])
        for j, (x, y, _c, cf) in enumerate(pts):
            conf[0, c, j], ys[0, c, j], xs[0, c, j] = cf, y, x
    return conf, ys, xs


@pytest.mark.parametrize("ci", [0, 1])
def test_pseudo_label_generation_golden(cl4, golden_more, ci):
    """train.py:451-466 + modules/utils.py:179-253 against reference outputs: everything bit-exact."""
    from cl4wsis_b200.modules.utils import pseudo_label_generation_batch
    g = golden_more("pseudo")
    k = f"pseudo_{ci}__"
    C = g[k + "center"].shape[0]
    conf, ys, xs = _peaks_from_points(g[k + "points"].tolist(), C, 4)
    c, o, w, m = pseudo_label_generation_batch(cuda(g[k + "gt"][None]), (cuda(conf), cuda(ys), cuda(xs)),
                                               cuda(g[k + "label"][None]), 0.7, int(g[k + "sigma"]))
    assert int(m[0]) == int(g[k + "match"])
    assert np.array_equal(c[0].cpu().numpy(), g[k + "center"])
    assert np.array_equal(o[0].cpu().numpy(), g[k + "offset"])
    assert np.array_equal(w[0].cpu().numpy(), g[k + "weight"])


def test_phase2_pseudo_labels_end_to_end_vs_oracle(cl4, oracle):
    """smoothing -> peak_extract -> pseudo_label_generation on the device for a batch, against the
    per-image oracle loop written as train.py:429-466 drives the reference."""
    from cl4wsis_b200.modules.utils import pseudo_label_generation_batch
    from cl4wsis_b200.wss.utils import peak_extract_device, smoothing
    rng = np.random.default_rng(99)
    B, C, H, W = 3, 20, 256, 320
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    gt = np.zeros((B, H, W), np.int64)
    cam = (0.05 * rng.random((B, C, H, W))).astype(np.float32)
    lab = np.zeros((B, C), np.float32)
    for b in range(B):
        for _ in range(10):
            cls = int(rng.integers(0, C)); cy, cx = int(rng.integers(10, H - 10)), int(rng.integers(10, W - 10))
            ry, rx = int(rng.integers(3, 40)), int(rng.integers(3, 40))
            gt[b][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = cls + 1
            lab[b, cls] = float(rng.random() < 0.8)
            for _p in range(int(rng.integers(0, 3))):  # 0, 1 or 2 CAM peaks per blob
                py, px = cy + int(rng.integers(-2, 3)), cx + int(rng.integers(-2, 3))
                cam[b, cls] = np.maximum(cam[b, cls], rng.uniform(0.6, 1.0) * np.exp(-((yy - py) ** 2 + (xx - px) ** 2) / 50).astype(np.float32))
    sm = smoothing(cuda(cam))
    peaks = peak_extract_device(sm, kernel=15, K=25)
    c, o, w, m = pseudo_label_generation_batch(cuda(gt), peaks, cuda(lab), 0.7, 6)
    sm_o = oracle.labelgen.smoothing(cam)
    np.testing.assert_allclose(sm.cpu().numpy(), sm_o, rtol=1e-6, atol=1e-7)
    conf, ys, xs = oracle.peak_extract(sm.cpu().numpy(), 15, 25)  # same smoothed map: isolates the label generation
    gsn = oracle.labelgen.gaussian(6)
    total = 0
    for b in range(B):
        pts = []
        for l in np.nonzero(lab[b])[0]:
            for cf, x, y in zip(conf[b, l], xs[b, l], ys[b, l]):
                if cf < np.float32(0.7):
                    break
                pts.append([x, y, l, cf])
        wc, wo, ww, n = oracle.labelgen.pseudo_label_generation(gt[b], pts, lab[b], C, 6, gsn)
        assert int(m[b]) == n
        total += n
        assert np.array_equal(c[b].cpu().numpy(), wc) and np.array_equal(o[b].cpu().numpy(), wo)
        assert np.array_equal(w[b].cpu().numpy(), ww)
    assert total >= 3


# --------------------------------------------------------------------------- validation path
class _VArgs:
    pass


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_get_ins_map_golden(cl4, golden_more, ci):
    """dataset/utils.py:795-902 (Trainer.validate, train.py:622) against reference outputs: same
    instances in the same order, masks and labels bit-exact, scores to fp32 rounding."""
    from cl4wsis_b200.dataset.utils import get_ins_map
    g = golden_more("insmap")
    k = f"insmap_{ci}__"
    a = _VArgs()
    a.val_thresh, a.val_kernel, a.beta, ign, flip, clean = g[k + "args"].tolist()
    a.val_kernel, a.val_ignore, a.val_flip, a.val_clean = int(a.val_kernel), bool(ign), bool(flip), bool(clean)
    out = {n: cuda(g[k + n]) for n in ("seg", "center", "offset")}
    tgt = tuple(int(v) for v in g[k + "target"])
    seg_map, pl, pm, ps = get_ins_map(out, cuda(g[k + "cls_label"]), tgt, torch.device("cuda"), a)
    assert np.array_equal(seg_map, g[k + "seg_map"])
    assert np.array_equal(pl, g[k + "pred_label"])
    shape = tuple(g[k + "pred_mask_shape"])
    want_mask = np.unpackbits(g[k + "pred_mask"], axis=-1)[..., :shape[-1]].astype(bool)
    assert pm.shape == shape and pm.dtype == np.bool_ and np.array_equal(pm, want_mask)
    np.testing.assert_allclose(ps, g[k + "pred_score"], rtol=2e-6, atol=0)
    assert np.array_equal(out["offset"].cpu().numpy(), g[k + "offset_after"])  # rescaled in place, as the reference


@pytest.mark.parametrize("H,W,n_inst,flip,clean,ignore,kern,thr,beta,noise", [
    (512, 512, 24, False, False, False, 41, 0.1, 3.0, 0.02),
    (512, 512, 30, True, True, True, 41, 0.1, 3.0, 0.02),
    (320, 480, 16, False, True, False, 7, 0.2, 5.0, 0.3),     # many centres per contour
    (200, 333, 12, True, False, False, 41, 0.1, 5.0, 0.09),   # clustered centres, odd width
])
def test_get_ins_map_random_vs_oracle(cl4, oracle, H, W, n_inst, flip, clean, ignore, kern, thr, beta, noise):
    """dataset/utils.py:795-902 at validation size against oracle/labelgen.get_ins_map (numpy + OpenCV, pinned to the
    reference on the insmap fixtures): seg map, instance order, labels and masks bit-exact, scores to fp32 rounding."""
    from cl4wsis_b200.dataset.utils import get_ins_map
    rng = np.random.default_rng(H * 7 + n_inst)
    C = 20
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    nb = 2 if flip else 1
    segs, heats, offs = [], [], []
    inst = [(int(rng.integers(0, C)), int(rng.integers(10, H - 10)), int(rng.integers(10, W - 10)),
             int(rng.integers(4, H // 6)), int(rng.integers(4, W // 6))) for _ in range(n_inst)]
    for v in range(nb):
        gt = np.zeros((H, W), np.int64)
        heat = (noise * rng.random((C, H, W))).astype(np.float32)
        off = (0.3 * rng.standard_normal((2, H, W))).astype(np.float32) + 40.0
        for i, (cls, cy, cx, ry, rx) in enumerate(inst):
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            gt[m] = cls + 1
            if i % 5 != 4:  # every fifth instance has no peak
                heat[cls] = np.maximum(heat[cls], rng.uniform(0.5, 0.95) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 72.0).astype(np.float32))
            off[0][m] = (cy - yy)[m] + 0.3 * rng.standard_normal(int(m.sum())).astype(np.float32)
            off[1][m] = (cx - xx)[m] + 0.3 * rng.standard_normal(int(m.sum())).astype(np.float32)
        seg = rng.standard_normal((C + 1, H, W)).astype(np.float32)
        for c in range(C + 1):
            seg[c][gt == c] += 3.0
        if v == 1:  # the second view is the mirrored image
            seg, heat, off = seg[..., ::-1].copy(), heat[..., ::-1].copy(), off[..., ::-1].copy()
        segs.append(seg); heats.append(heat); offs.append(off)
    o = {"seg": np.stack(segs), "center": np.stack(heats), "offset": np.stack(offs)}
    cls_label = np.zeros((1, C), np.float32)
    for it in inst[:-2]:
        cls_label[0, it[0]] = 1
    tgt = (H * 2, W + 17)
    a = _VArgs()
    a.val_thresh, a.val_kernel, a.beta, a.val_ignore, a.val_flip, a.val_clean = thr, kern, beta, ignore, flip, clean
    w_seg, w_lab, w_mask, w_score, w_off = oracle.labelgen.get_ins_map(o["seg"], o["center"], o["offset"], cls_label, tgt, thr,
                                                                        kern, beta, ignore, flip, clean)
    out = {n: cuda(v) for n, v in o.items()}
    seg_map, pl, pm, ps = get_ins_map(out, cuda(cls_label), tgt, torch.device("cuda"), a)
    assert np.array_equal(seg_map, w_seg)
    assert np.array_equal(pl, w_lab), (pl.tolist(), w_lab.tolist())
    assert pm.dtype == np.bool_ and pm.shape == w_mask.shape and np.array_equal(pm, w_mask)
    np.testing.assert_allclose(ps, w_score, rtol=2e-6, atol=0)
    assert len(pl) > n_inst // 3
    assert np.array_equal(out["offset"][0].cpu().numpy(), w_off)


# --------------------------------------------------------------------------- phase-1 producers / consumers of PAMR
@pytest.mark.parametrize("name", ["p1_one", "p1_rect", "p1_same", "p1_voc"])
def test_phase1_pieces_match_reference(cl4, golden_more, name):
    """denorm, denorm + bilinear shrink, class softmax, label gating + pseudo_gtmask (train.py:372-385) against the
    reference's own outputs; thresholded outputs bit-exact on identical inputs."""
    from cl4wsis_b200.utils import denorm
    from cl4wsis_b200.wss import single_stage as ss
    g = golden_more("phase1")
    k = lambda s: g[f"{name}__{s}"]  # noqa: E731
    assert np.array_equal(denorm(cuda(k("img"))).cpu().numpy(), k("denorm"))
    im = ss.denorm_resize(cuda(k("img")), k("im").shape[-2:]).cpu().numpy()
    np.testing.assert_allclose(im, k("im"), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(ss.softmax_channels(cuda(k("logits"))).cpu().numpy(), k("soft"), rtol=2e-6, atol=1e-9)
    assert np.array_equal(ss.pseudo_gtmask(cuda(k("soft")), ambiguous=False).cpu().numpy(), k("pseudo_noamb"))
    # gating + thresholds on the reference's PAMR output: bit-exact
    lib, L = cl4._lib.load(), cl4._lib
    pam = cuda(k("pamr").copy())
    B, C, h, w = pam.shape
    pseudo, thr = torch.empty_like(pam), torch.empty((B, C), device="cuda")
    L.check(lib.cl4_pseudo_gtmask(L.ptr(pam), L.ptr(cuda(k("l1h"))), L.ptr(pam), L.ptr(pseudo), L.ptr(thr), B, C, h * w,
                                  0.6, 0.7, 0.2, 1, L.stream_ptr()), "pseudo_gtmask")
    assert np.array_equal(pam.cpu().numpy(), k("gated"))
    assert np.array_equal(pseudo.cpu().numpy(), k("pseudo"))


@pytest.mark.parametrize("name", ["p1_one", "p1_rect", "p1_same", "p1_voc"])
def test_phase1_fused_step(cl4, golden_more, name):
    """The whole of train.py:372-385 through phase1_pseudo_labels (softmax, denorm + shrink, PAMR with the trainer's
    five dilations, gating, pseudo_gtmask).  The refined masks match to PAMR's tolerance; a pseudo-label may only
    differ where the reference's own mask sits within that tolerance of its threshold."""
    from cl4wsis_b200.wss import single_stage as ss
    g = golden_more("phase1")
    k = lambda s: g[f"{name}__{s}"]  # noqa: E731
    pamr = cl4.PAMR(num_iter=10, dilations=[1, 2, 4, 8, 12]).cuda()
    soft, pseudo = ss.phase1_pseudo_labels(cuda(k("img")), cuda(k("logits")), cuda(k("l1h")), pamr)
    np.testing.assert_allclose(soft.cpu().numpy(), k("gated"), rtol=RTOL, atol=ATOL)
    ref = k("gated")
    B, C = ref.shape[:2]
    mx = ref.reshape(B, C, -1).max(-1)
    thr = np.maximum(mx * np.where(np.arange(C) == 0, np.float32(0.7), np.float32(0.6))[None, :], np.float32(0.2))[:, :, None, None]
    near = np.abs(ref - thr) <= 2e-4 * np.abs(thr) + 2e-6
    near_px = near.any(axis=1, keepdims=True)  # the ambiguity rule couples the classes of a pixel
    diff = pseudo.cpu().numpy() != k("pseudo")
    assert not (diff & ~near_px).any(), int((diff & ~near_px).sum())


@pytest.mark.parametrize("B,C,h,w,Hi,Wi,dil", [
    (16, 21, 32, 32, 512, 512, [1, 2, 4, 8, 12]),      # the VOC trainer's shapes: 7 CTAs per image share the ambiguity rule
    (4, 81, 56, 56, 448, 448, [1, 2, 4, 8, 12]),       # coco-voc: 2 x 2 tiles, 41 CTAs per image
    (3, 5, 19, 26, 75, 102, [1, 2, 4, 8, 12, 24]),     # odd sizes, six dilations
    (2, 2, 64, 64, 64, 64, [1, 3]),                    # shrink factor 1, runtime dilation set
])
def test_phase1_two_launch_path_equals_separate_kernels(cl4, B, C, h, w, Hi, Wi, dil):
    """cl4_phase1_pseudo_labels (2 launches: prologue + on-chip PAMR with the gating / pseudo_gtmask epilogue) against the
    same step run as separate kernels (train.py:372-385): identical pseudo labels, soft masks equal to rounding."""
    from cl4wsis_b200.wss import single_stage as ss
    g = torch.Generator(device="cpu").manual_seed(B * 100 + C)
    images = ((torch.rand((B, 3, Hi, Wi), generator=g) - 0.45) / 0.226).cuda()
    logits = (3 * torch.randn((B, C, h, w), generator=g)).cuda()
    l1h = (torch.rand((B, C - 1), generator=g) < 0.4).float().cuda()
    l1h[:, 0] = 1
    mod = cl4.PAMR(10, dil).cuda()
    assert ss._fused_applicable(mod, h, w) and ss.PHASE1_LAUNCHES == 2
    soft, pseudo = ss.phase1_pseudo_labels(images, logits, l1h, mod)
    soft_u, pseudo_u = ss.phase1_pseudo_labels_unfused(images, logits, l1h, mod)
    torch.testing.assert_close(soft, soft_u, rtol=1e-6, atol=1e-8)
    if torch.equal(soft, soft_u):
        assert torch.equal(pseudo, pseudo_u)
    else:  # a pseudo label may only differ where the mask sits within rounding of its threshold
        assert (pseudo != pseudo_u).float().mean().item() < 1e-4
    assert pseudo.sum(1).max().item() <= 1.0           # ambiguous pixels are cleared across ALL classes of an image
    assert 0 < pseudo.sum().item() < pseudo.numel()


def test_phase1_error_behaviour(cl4):
    from cl4wsis_b200.utils import denorm
    with pytest.raises(AssertionError):
        denorm(torch.zeros(2, 4, 8, 8, device="cuda"))           # utils/utils.py:35 "Expected RGB image"
    with pytest.raises(RuntimeError):
        denorm(torch.zeros(2, 3, 8, 8))                           # CPU tensor: no fallback


# --------------------------------------------------------------------------- second oracle: stock PyTorch ON THE GPU
@pytest.fixture()
def fp32_convs():
    """ATen's CUDA convolutions default to TF32 (torch.backends.cudnn.allow_tf32); the reference's one-hot shift kernels
    would then see masks rounded to 10 mantissa bits.  The parity bar is fp32 arithmetic, so TF32 is off here."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("B,C,H,W,dil,T", [
    (2, 21, 128, 160, [1, 2, 4, 8, 12, 24], 10),   # class-pair sweep, odd class count
    (1, 4, 256, 256, [1, 2, 4, 8, 12], 10),        # the trainer's dilation set
    (16, 21, 32, 32, [1, 2, 4, 8, 12], 10),        # the trainer's real regime (fused small-map kernel), train.py:376-379
    (2, 3, 56, 56, [1, 2, 4, 8, 12], 10),
])
def test_pamr_vs_stock_pytorch_on_cuda(cl4, fp32_convs, B, C, H, W, dil, T):
    """wss/modules.py:133-152 executed by ATen's CUDA kernels (F.pad + F.conv2d + std + softmax) on the same device."""
    from oracle import torch_ref as tr
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + C)
    x = (torch.randint(0, 256, (B, 3, H, W), generator=g).float() / 255).cuda()
    m = torch.randn((B, C, H, W), generator=g).softmax(1).cuda()
    want = tr.pamr(x, m, T, dil)
    got = cl4.PAMR(T, dil).cuda()(x, m)
    torch.testing.assert_close(got, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("H,W,Kc", [(512, 512, 5), (256, 320, 200), (64, 50, 1000), (1024, 1024, 50)])
def test_group_pixels_vs_torch_norm_on_cuda(cl4, H, W, Kc):
    """SURVEY §7.2 / ADVICE: the ids are pinned to ATen's CPU 2-vector norm, sqrt(fma(dx,dx,rn(dy^2))).  ATen's CUDA norm
    kernel may round differently; the trainer runs group_pixels on CUDA (train.py:492 -> modules/utils.py:604).
    Requirement: the ids equal torch.norm + argmin ON THIS DEVICE except where the two nearest centres are tied to
    within 2 ulp of the distance (there the argmin depends on one rounding, and the CPU reference is the pin)."""
    from oracle import torch_ref as tr
    from cl4wsis_b200.modules import utils as mu
    g = torch.Generator(device="cpu").manual_seed(H + Kc)
    ctr = torch.stack([torch.randint(0, H, (Kc,), generator=g), torch.randint(0, W, (Kc,), generator=g)], 1).cuda()
    off = (20 * torch.randn((1, 2, H, W), generator=g)).cuda()
    got = mu.group_pixels(ctr, off)
    want = tr.group_pixels(ctr.float(), off)
    diff = got != want
    n_bad = int(diff.sum())
    if n_bad:
        yy, xx = torch.nonzero(diff[0], as_tuple=True)
        loc = torch.stack([yy.float() + off[0, 0, yy, xx], xx.float() + off[0, 1, yy, xx]], 1).double()
        c = ctr.double()
        d_got = (c[got[0, yy, xx] - 1] - loc).norm(dim=-1)
        d_want = (c[want[0, yy, xx] - 1] - loc).norm(dim=-1)
        rel = ((d_got - d_want).abs() / d_want.clamp_min(1e-30)).max().item()
        assert rel <= 2.5e-7, f"{n_bad} ids differ from torch-CUDA and are not fp32 ties (rel distance gap {rel:.3e})"
    assert n_bad <= max(4, H * W * Kc // 200000), f"{n_bad} tie-order differences vs torch-CUDA"


@pytest.mark.parametrize("H,W,k,thr", [(512, 512, 41, 0.3), (100, 333, 3, 0.1), (65, 31, 7, 0.5)])
def test_center_nms_vs_torch_maxpool_on_cuda(cl4, H, W, k, thr):
    """modules/utils.py:478-492 on CUDA tensors: F.threshold + F.max_pool2d + nonzero -> identical centre list."""
    from oracle import torch_ref as tr
    from cl4wsis_b200.modules import utils as mu
    g = torch.Generator(device="cpu").manual_seed(H * 3 + k)
    heat = (torch.rand((1, 1, H, W), generator=g) * 64).round() / 64   # exact plateaus
    heat = heat.cuda()
    want = tr.find_instance_center(heat.clone(), thr, k)
    got = mu.find_instance_center(heat.clone(), thr, k)
    assert torch.equal(got, want)


def test_peak_extract_and_smoothing_vs_torch_on_cuda(cl4):
    from oracle import torch_ref as tr
    from cl4wsis_b200.wss.utils import peak_extract_device, smoothing
    g = torch.Generator(device="cpu").manual_seed(99)
    heat = torch.rand((2, 20, 128, 96), generator=g).cuda()
    sm = smoothing(heat, 3)
    torch.testing.assert_close(sm, tr.smoothing(heat, 3), rtol=2e-6, atol=1e-7)
    s, y, x = peak_extract_device(sm, 15, 25)
    ws, wy, wx = tr.peak_extract(sm, 15, 25)
    assert torch.equal(s, ws)                      # scores exact; random floats: no ties, so indices too
    assert torch.equal(y, wy) and torch.equal(x, wx)


# --------------------------------------------------------------------------- same images, sharded two ways
def test_sharding_reproduces_the_single_gpu_result(cl4):
    """images are independent, so a batch split across ranks must give exactly what one GPU gives.
    One GPU plays both ranks: images 0..7 as one batch of 8 vs two shards of 4 (bench.py seeds synthetic images by GLOBAL
    image index); refined masks, ids and centre counts must be bit-identical and the checksums equal."""
    import bench
    cfg = dict(B=8, C=21, H=128, W=128, dil=[1, 2, 4, 8, 12, 24], T=10, Kc=5, nms=41, thr=0.3)
    full = [t.cuda() for t in bench.synth_inputs(cfg, first_image=0, n_images=8)]
    step8 = cl4.PseudoLabelStep(8, 21, 128, 128, num_iter=10, threshold=0.3, nms_kernel=41)
    r8, i8, c8, _ = (t.clone() for t in step8.run(*full))
    step4 = cl4.PseudoLabelStep(4, 21, 128, 128, num_iter=10, threshold=0.3, nms_kernel=41)
    ck = 0.0
    for rank in range(2):
        shard = [t.cuda() for t in bench.synth_inputs(cfg, first_image=4 * rank, n_images=4)]
        for a, b in zip(shard, full):
            assert torch.equal(a, b[4 * rank:4 * rank + 4])
        r4, i4, c4, _ = step4.run(*shard)
        assert torch.equal(r4, r8[4 * rank:4 * rank + 4])
        assert torch.equal(i4, i8[4 * rank:4 * rank + 4])
        assert torch.equal(c4, c8[4 * rank:4 * rank + 4])
        ck += float(i4.double().sum())
    assert ck == float(i8.double().sum())
