"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean and share per kernel."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1:][skip:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("cl4::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)  # -> us
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'mean us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:5d} {t / n:10.1f} {100 * t / tot:6.1f}%")
print(f"total {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
