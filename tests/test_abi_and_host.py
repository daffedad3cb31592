"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the header
declares (no compute calls), the ctypes table matches the header, and host-side argument
validation mirrors the reference's error behaviour without touching a GPU."""
import ctypes

import numpy as np
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cl4wsis_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cl4_[a-z0-9_]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def lib():
    from cl4wsis_b200 import build
    build.build()
    from cl4wsis_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cl4wsis_b200.h but not exported"


def test_ctypes_table_matches_header():
    from cl4wsis_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_abi_version_and_scratch_queries(lib):
    assert lib.cl4_abi_version() == 1
    # pure host arithmetic, no CUDA call: weights [B,48,H,W] + two replicate-padded mask buffers + the padded image
    n = lib.cl4_pamr_scratch_bytes(16, 3, 21, 512, 512, 6, 10)
    assert n == 4 * 16 * 512 * 512 * 48 + 2 * 4 * 16 * 21 * 560 * 560 + 4 * 16 * 3 * 560 * 560
    assert lib.cl4_pamr_scratch_bytes(16, 3, 21, 512, 512, 6, 1) == 4 * 16 * 512 * 512 * 48 + 4 * 16 * (21 + 3) * 560 * 560
    assert lib.cl4_center_nms_scratch_bytes(1, 512, 512) >= 512 * 16 * 4 + 512 * 4
    assert lib.cl4_peak_extract_scratch_bytes(2, 3, 64, 64, 15, 25) == (2 * 3 * 2 * 25 + 2 * 3) * 8  # candidates + one bound key per plane
    assert lib.cl4_peak_extract_scratch_bytes(1, 1, 64, 64, 15, 1000) == (2 * 256 + 1) * 8     # K > 256: rounds of 256


def test_argument_validation_returns_codes_without_a_gpu(lib):
    from cl4wsis_b200 import _lib
    null = ctypes.c_void_p(0)
    dil = _lib.int_array([1, 2])
    assert lib.cl4_pamr_sweep(null, null, null, 1, 0, 8, 8, dil, 2, null) == _lib.CL4_EINVAL
    assert lib.cl4_pamr_weights(null, null, 1, 3, 8, 8, dil, 9, null) == _lib.CL4_EUNSUPPORTED
    assert b"dilations" in lib.cl4_last_error()
    assert lib.cl4_pamr_weights(null, null, 1, 3, 8, 8, _lib.int_array([1, 0]), 2, null) == _lib.CL4_EINVAL
    assert lib.cl4_center_nms(null, 0.1, 0.0, 4, 1, 8, 8, null, null, 0, null, 0, null) == _lib.CL4_EINVAL
    assert b"odd" in lib.cl4_last_error()
    assert lib.cl4_peak_extract(null, null, null, null, null, 0, 1, 1, 8, 8, 3, 65, null) == _lib.CL4_EINVAL
    assert lib.cl4_peak_extract(null, null, null, null, null, 0, 1, 1, 64, 64, 3, 300, null) == _lib.CL4_EINVAL  # K is fine; null pointers
    assert b"null" in lib.cl4_last_error()
    assert lib.cl4_group_pixels(null, null, 1, 1, null, null, null, 1, 8, 8, 0, null) == _lib.CL4_EINVAL
    # misaligned buffers are refused before anything is launched (128-bit loads, TMA)
    v = ctypes.c_void_p
    assert lib.cl4_pamr_forward(v(0x1004), v(0x2000), v(0x3000), v(0x4000), 1 << 40, 1, 3, 2, 64, 64, _lib.int_array([1, 2, 4, 8, 12, 24]),
                                6, 10, null) == _lib.CL4_EINVAL
    assert b"16-byte aligned" in lib.cl4_last_error()
    with pytest.raises(ValueError):
        _lib.check(_lib.CL4_EINVAL, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.CL4_EUNSUPPORTED, "x")


def test_no_cpu_fallback_and_reference_error_conventions():
    import cl4wsis_b200 as cl4
    with pytest.raises(RuntimeError, match="CUDA"):
        cl4.PAMR()(torch.rand(1, 3, 8, 8), torch.rand(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        cl4.peak_extract(torch.rand(1, 1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        cl4.find_instance_center(torch.rand(1, 1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        cl4.group_pixels(torch.zeros(1, 2, dtype=torch.long), torch.rand(1, 2, 8, 8))


def test_pamr_module_mirrors_reference_state():
    import cl4wsis_b200 as cl4
    mod = cl4.PAMR(num_iter=10, dilations=[1, 2, 4, 8, 12])
    assert mod.num_iter == 10 and mod.aff_x.dilations == [1, 2, 4, 8, 12]
    sd = mod.state_dict()
    assert sorted(sd) == ["aff_m.kernel", "aff_std.kernel", "aff_x.kernel"]
    k = sd["aff_x.kernel"]
    assert float(k[0, 0, 1, 1]) == 1 and float(k[0, 0, 0, 0]) == -1 and float(k[4, 0, 1, 2]) == -1
    assert float(sd["aff_m.kernel"].sum()) == 8 and float(sd["aff_std.kernel"].sum()) == 9
    assert len(list(mod.parameters())) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cl4wsis_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "cl4_oracle" not in src, f


def test_opencv_8_connectivity_label_order_rule():
    """Host logic of the validation twin: the order in which OpenCV numbers 8-connected components
    (2x2-block raster scan) reproduced from (first pixel, first block column) — checked against cv2."""
    import cv2
    from cl4wsis_b200.dataset.utils import _opencv_label_order
    rng = np.random.default_rng(0)
    for _ in range(100):
        H, W = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        m = (rng.random((H, W)) < rng.uniform(0.2, 0.7)).astype(np.uint8)
        n, lab, _, _ = cv2.connectedComponentsWithStats(m, connectivity=8)
        # hand the function a scrambled slot numbering, as the GPU's atomically assigned slots are
        perm = rng.permutation(n - 1)
        comp = np.full((H, W), -1, np.int32)
        info = np.zeros((n - 1, 5), np.int32)
        for k in range(1, n):
            comp[lab == k] = perm[k - 1]
            info[perm[k - 1], 0] = np.flatnonzero((lab == k).ravel())[0]
        order = _opencv_label_order(comp, info, list(range(n - 1)))
        assert order == [int(perm[k - 1]) for k in range(1, n)]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract: exactly one JSON line on stdout; the reference arm runs on the CPU (oracle port) with the
    base contract's keys plus impl / cpu_baseline / e2e."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-images", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms print the SAME config object (the driver compares them key for key)
    import bench
    assert d["config"] == bench.bench_config("voc_b16_c21_512", bench.WORKLOADS["voc_b16_c21_512"])


def test_bench_inputs_do_not_depend_on_sharding():
    """Synthetic images are seeded by GLOBAL image index: rank r's batch is images r*B .. r*B+B-1 of the job."""
    import bench
    import torch
    cfg = dict(B=4, C=3, H=32, W=32, dil=[1, 2], T=1, Kc=2, nms=5, thr=0.3)
    whole = bench.synth_inputs(cfg, 0, 4)
    for r in range(2):
        part = bench.synth_inputs(cfg, 2 * r, 2)
        for a, b in zip(part, whole):
            assert torch.equal(a, b[2 * r:2 * r + 2])


def test_traffic_lookup_refuses_stale_captures(tmp_path, monkeypatch):
    """roofline.traffic is read from profiles/traffic.json only for a capture of THIS build of the sweep kernels."""
    import json
    import bench
    sha = bench.sweep_sources_sha()
    doc = {"entries": [
        {"workload": "w1", "kernel": "k", "sources_sha": sha, "dram_bytes_read": 3.0, "dram_bytes_write": 4.0, "report": "a"},
        {"workload": "w2", "kernel": "k", "sources_sha": "0" * 16, "dram_bytes_read": 3.0, "dram_bytes_write": 4.0, "report": "b"}]}
    (tmp_path / "profiles").mkdir()
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps(doc))
    (tmp_path / "cl4wsis_b200").symlink_to(os.path.join(ROOT, "cl4wsis_b200"))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_traffic("w1", "k")[0] == 7.0
    assert bench.measured_traffic("w2", "k")[0] is None and "this build" in bench.measured_traffic("w2", "k")[1]
    assert bench.measured_traffic("w3", "k")[0] is None


def test_lattice_sweep_layout_invariants(lib):
    """The (pixel -> thread, slot) maps of the default propagation kernel (pamr_lattice.cu), read through the C ABI:
    a bijection per warp group; A's warp j and B's warp j own the same 8 tile rows (their partial sums meet per warp pair);
    the lane patterns are bank-conflict free at the window pitch of 84 floats and the partial-sum pitch of 36."""
    import ctypes
    t, s = ctypes.c_int(), ctypes.c_int()
    own = {}
    for g in (0, 1):
        seen = set()
        for y in range(32):
            for x in range(32):
                assert lib.cl4_lattice_owner(g, y, x, ctypes.byref(t), ctypes.byref(s)) == 0
                assert 0 <= t.value < 128 and 0 <= s.value < 8
                seen.add((t.value, s.value))
                own[g, y, x] = (t.value, s.value)
        assert len(seen) == 1024
    assert lib.cl4_lattice_owner(2, 0, 0, ctypes.byref(t), ctypes.byref(s)) != 0
    assert lib.cl4_lattice_owner(0, 32, 0, ctypes.byref(t), ctypes.byref(s)) != 0
    for g in (0, 1):  # warp j <-> rows 8j .. 8j+7
        for (gg, y, x), (th, _) in own.items():
            if gg == g:
                assert th // 32 == y // 8
    # block origin (slot 0) of every thread
    origin = {(g, th): (y, x) for (g, y, x), (th, sl) in own.items() if sl == 0}
    # group A: slots are a 2 x 4 lattice block of spacing 4
    for (g, y, x), (th, sl) in own.items():
        if g == 0:
            oy, ox = origin[0, th]
            assert (y - oy, x - ox) == (4 * (sl // 4), 4 * (sl % 4))
        else:  # group B: 4 x 2 block of adjacent pixels
            oy, ox = origin[1, th]
            assert (y - oy, x - ox) == (sl // 2, sl % 2)
    for pitch in (84, 36):
        for w in range(4):
            # A: 32-bit accesses, all 32 lanes of a warp in distinct banks
            banks = {(origin[0, 32 * w + ln][0] * pitch + origin[0, 32 * w + ln][1]) % 32 for ln in range(32)}
            assert len(banks) == 32, (pitch, w)
            # B: 64-bit accesses, the 16 lanes of a half-warp in distinct bank pairs
            for h in range(2):
                pairs = {((origin[1, 32 * w + 16 * h + ln][0] * pitch + origin[1, 32 * w + 16 * h + ln][1]) // 2) % 16 for ln in range(16)}
                assert len(pairs) == 16, (pitch, w, h)
