import csv, sys, collections, re, subprocess
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__cycles_active.avg','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','lts__t_bytes.sum']
d = data[0]
print('---', d[hdr.index('Kernel Name')][:70])
for k in keys:
    if k in hdr: print(f"{k:75s} {d[hdr.index(k)]:>20s} {units[hdr.index(k)]}")
for i,h in enumerate(hdr):
    if 'stalled' in h and 'ratio' in h and 'not_issued' not in h:
        try: v=float(d[i])
        except: continue
        if v>0.1: print(f"  {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.3f}")
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi+1:] if len(r)==len(hdr) and r[0]!="Address"]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
op = collections.Counter(); samp = collections.Counter(); tot=0; tots=0
for r in data:
    try: n=int(r[iexe]); s_=int(r[isamp])
    except: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    o = m.group(2).split('.')[0] if m else '?'
    op[o]+=n; samp[o]+=s_; tot+=n; tots+=s_
print("total inst", tot, "samples", tots)
for o,n in op.most_common(14): print(f"{o:12s} {n:12d} {100*n/tot:5.1f}%  samples {samp[o]:7d} {100*samp[o]/tots:5.1f}%")
def key(r):
    try: return -int(r[isamp])
    except: return 0
for r in sorted(data, key=key)[:int(sys.argv[2]) if len(sys.argv)>2 else 25]: print(r[ia], r[isamp], r[iexe], r[isrc][:100])
