"""Quick text summary of an ncu capture: key metrics, stall reasons, opcode mix per warp-plane, top stall sites.
usage: python scripts/ncu_quick.py <prefix>   (reads <prefix>_raw.csv and <prefix>_src.csv made with ncu -i ... --page raw|source --csv)"""
import collections, csv, re, sys
pre = sys.argv[1]
rows = list(csv.reader(open(pre + "_raw.csv")))
hdr = rows[0]; r = rows[2] if len(rows) > 2 else rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__waves_per_multiprocessor', 'sass__inst_executed_local_loads', 'sass__inst_executed_shared_loads',
        'sass__inst_executed_shared_stores', 'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print("%-70s %s %s" % (w, r[i], rows[1][i]))
print("stalls per issue:", ", ".join("%s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[i]))
      for i, h in enumerate(hdr) if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(r[i] or 0) > 0.03))
rows = list(csv.reader(open(pre + "_src.csv")))
hdr = rows[1]; ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); iss = hdr.index('Warp Stall Sampling (All Samples)')
R = [x for x in rows[2:] if len(x) > iss]
c = collections.Counter(); tot = 0
for x in R:
    s = re.sub(r'^@!?U?P\d+\s+', '', x[ia].strip()); n = int(x[ie] or 0)
    c[s.split()[0] if s else ''] += n; tot += n
bars = max(1, c['BAR.SYNC.DEFER_BLOCKING'] / 2)
print("instr per warp-step: %.1f" % (tot / bars))
print("  ".join("%s %.1f" % (k, v / bars) for k, v in c.most_common(30)))
st = sorted(((int(x[iss] or 0), x[ia].strip()[:60]) for x in R), reverse=True)
T = sum(x[0] for x in st)
print("top stall sites of %d samples:" % T)
for x in st[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]: print("  %5d %s" % x)
