"""One block per kernel of an `ncu --set full` report: duration, DRAM traffic, unit utilisations, issue rate, top stalls.
Usage: python tools/ncu_multi_summary.py rep.ncu-rep > profiles/rXX_secondary_ncu_summary.txt"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor']
seen = {}
for d in data:
    name = d[hdr.index('Kernel Name')]
    short = name.split('(')[0][-60:]
    seen[short] = seen.get(short, 0) + 1
    if seen[short] > 2:
        continue
    print('---', name[:110])
    for k in keys:
        if k in hdr:
            print(f"  {k:72s} {d[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(d[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('  stalls (warps per issue): ' + ', '.join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))
