"""Record the measured DRAM traffic of the sweep kernels from an `ncu --set full` report into profiles/traffic.json.

    python tools/record_traffic.py gpurun_out/r02f_duo.ncu-rep --workload voc_b16_c21_512 [--copy-to profiles/r02f_sweep]

bench.py prints `roofline.traffic` from that file, and only when the content hash of the sweep kernels' sources
(bench.sweep_sources_sha) equals the hash recorded here -- a capture of an older build reads as null, never as a number.
With --copy-to the raw CSV of the report is also written next to the summary (profiles/ is tracked; gpurun_out/ is not).
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
KERNELS = ("pamr_sweep_duo_kernel", "pamr_sweep_lattice_kernel", "pamr_sweep_tma_kernel")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--copy-to", default=None, help="prefix under profiles/ for the raw CSV")
    a = ap.parse_args()
    import bench
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        return float(r[ix[key]].replace(",", "")) * SCALE.get(units[ix[key]], 1.0)

    per = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        k = next((k for k in KERNELS if k in name), None)
        if k is None:
            continue
        per.setdefault(k, []).append((val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")))
    if not per:
        raise SystemExit("no sweep kernel in the report")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            doc = json.load(f)
    except FileNotFoundError:
        doc = {"comment": "ncu --set full, dram__bytes_{read,write}.sum per launch; written by tools/record_traffic.py", "entries": []}
    sha = bench.sweep_sources_sha()
    for k, v in per.items():
        n = len(v)
        e = {"workload": a.workload, "kernel": k, "sources_sha": sha, "launches_captured": n,
             "dram_bytes_read": sum(x[0] for x in v) / n, "dram_bytes_write": sum(x[1] for x in v) / n,
             "duration_us_under_ncu": sum(x[2] for x in v) / n, "report": os.path.basename(a.report)}
        doc["entries"] = [o for o in doc["entries"] if not (o["workload"] == a.workload and o["kernel"] == k)] + [e]
        print(e)
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    if a.copy_to:
        with open(os.path.join(ROOT, a.copy_to + "_ncu_raw.csv"), "w") as f:
            f.write(raw)


if __name__ == "__main__":
    main()
