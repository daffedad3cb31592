import csv, collections, re, sys, subprocess
src = subprocess.run(['ncu','-i',sys.argv[1],'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr = rows[hi]; data=[r for r in rows[hi+1:] if len(r)==len(hdr) and r[0]!="Address"]
isrc, isamp, iexe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
groups = collections.defaultdict(collections.Counter); gs=collections.Counter()
for r in data:
    try: n=int(r[iexe]); s=int(r[isamp])
    except: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc]); o = m.group(2).split('.')[0] if m else '?'
    groups[n][o]+=1; gs[n]+=s
for n,c in sorted(groups.items(), key=lambda kv:-gs[kv[0]])[:6]:
    tot=sum(c.values()); print(f"exec={n} ninstr={tot} total={n*tot/1e6:.1f}M samples={gs[n]}", c.most_common(6))
def key(r):
    try: return -int(r[isamp])
    except: return 0
for r in sorted(data, key=key)[:8]: print(r[isamp], r[iexe], r[isrc][:90])
