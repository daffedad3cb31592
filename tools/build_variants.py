"""Build A/B variants of the library: python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 ...  -> cl4wsis_b200/libcl4_<name>.so
Only the files named in CL4_VARIANT_SOURCES (default pamr_lattice.cu) are recompiled with the flags."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cl4wsis_b200 import build as b
b.build()
srcs = os.environ.get("CL4_VARIANT_SOURCES", "pamr_lattice.cu").split(",")
procs = []
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    flags = [f for f in flags.split(",") if f]
    objs = []
    for src in b.SOURCES:
        o = os.path.join(b.HERE, "build", src.replace(".cu", ".o"))
        if src in srcs:
            o = os.path.join(b.HERE, "build", f"{name}_{src.replace('.cu', '.o')}")
            procs.append((name, subprocess.Popen([b._nvcc(), "-c", os.path.join(b.CSRC, src), "-o", o] + b.NVCC_FLAGS + flags)))
        objs.append(o)
    procs.append((name, objs))
for name, p in procs:
    if isinstance(p, subprocess.Popen):
        assert p.wait() == 0, name
for name, p in procs:
    if isinstance(p, list):
        out = os.path.join(b.HERE, f"libcl4_{name}.so")
        subprocess.check_call([b._nvcc(), "-shared", "-o", out] + p + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
        print(out)
