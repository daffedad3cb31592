"""Builds ``libcl4wsis_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m cl4wsis_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcl4wsis_b200.so")
SOURCES = ["core.cu", "group.cu", "nms.cu", "pamr.cu", "pamr_tma.cu", "pamr_lattice.cu", "pamr_duo.cu", "pamr_fused.cu", "ccl.cu", "stencil.cu", "refine.cu", "phase1.cu", "phase1_fused.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=None, lib=None):
    """``extra``: additional nvcc flags (A/B experiments, e.g. ["-DCL4_SWEEP_STAGES=6"]);
    ``lib``: alternative output path for such a variant (objects are rebuilt)."""
    extra = list(extra or []) + os.environ.get("CL4_NVCC_EXTRA", "").split()
    if extra:
        force = True
    out_lib = lib or LIB
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "cl4wsis_b200.h"))
    objdir = os.path.join(HERE, "build" if not lib else "build_variant")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, "-c", s, "-o", o] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else [])
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = force or bool(procs) or _stale(out_lib, objs)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose or "warning" in out:
            sys.stderr.write(out)
    if relink:
        cmd = [nvcc, "-shared", "-o", out_lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.check_call(cmd)
    return out_lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
