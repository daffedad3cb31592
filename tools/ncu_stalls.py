"""Per-instruction stall breakdown from an ncu report's source page, grouped into code regions between mbarrier waits.
Usage: python tools/ncu_stalls.py rep.ncu-rep [lo_index hi_index]"""
import csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def num(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
tot = sum(num(r, '# Samples') for r in data)
print("total samples", tot, "instructions", len(data))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi_ = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
# region summary: split at every SYNCS.PHASECHK (a wait) ; print per region: samples, executed of first instr, top stalls
reg_start = lo
def flush(a, b):
    if b <= a: return
    s = sum(num(r, '# Samples') for r in data[a:b])
    if s < tot * 0.004: return
    agg = {k: sum(num(r, k) for r in data[a:b]) for k in st}
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:5]
    ex = max(num(r, 'Instructions Executed') for r in data[a:b])
    n_ffma = sum(1 for r in data[a:b] if 'FFMA' in r[ix['Source']])
    n_lds = sum(1 for r in data[a:b] if 'LDS' in r[ix['Source']])
    print(f"[{a:5d},{b:5d}) samples {s:6d} {100*s/tot:5.1f}%  max-exec {ex:8d} ffma {n_ffma:4d} lds {n_lds:4d}  " + " ".join(f"{k[6:]}={v}" for k, v in top if v))
for i in range(lo, hi_):
    if 'SYNCS.PHASECHK' in data[i][ix['Source']]:
        flush(reg_start, i)
        reg_start = i
flush(reg_start, hi_)
if len(sys.argv) > 4:
    for i in range(lo, hi_):
        r = data[i]
        s = num(r, '# Samples')
        tops = sorted(((num(r, k), k[6:]) for k in st), reverse=True)[:3]
        print(i, r[0][-5:], s, num(r, 'Instructions Executed'), r[ix['Source']][:70], " ".join(f"{k}={v}" for v, k in tops if v))
