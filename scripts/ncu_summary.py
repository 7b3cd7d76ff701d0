"""Summarise ncu outputs brought back in gpurun_out/ into markdown for profiles/ (runs without a GPU).

usage: python scripts/ncu_summary.py launches <launches.csv> <launches-per-episode>
       python scripts/ncu_summary.py full <prof.ncu-rep>
"""
import collections
import csv
import subprocess
import sys


def launches(path, per_episode):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ep = rows[hi + 1:][-per_episode:]
    tot, cnt = collections.Counter(), collections.Counter()
    for r in ep:
        name = r[iK].split("(")[0].replace("void ", "").replace("mgpu::", "").replace("<unnamed>::", "")
        v = float(r[iV].replace(",", ""))
        ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[r[iU]] * v
        tot[name] += ms
        cnt[name] += 1
    T = sum(tot.values())
    print("last episode: %d launches, %.3f ms summed kernel time (ncu: serialised, cold-cache)\n" % (len(ep), T))
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in tot.most_common():
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v, 100 * v / T))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    iN = hdr.index("Kernel Name")
    for vals in rows[2:]:
        print("### `%s`\n" % vals[iN])
        print("| metric | value | unit |\n|---|---|---|")
        d = dict(zip(hdr, zip(vals, units)))
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        print()
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
    ops, tot = collections.Counter(), 0
    for r in rows[2:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        parts = r[iS].split()
        if parts and parts[0].startswith("@"):
            parts = parts[1:]
        if not parts:
            continue
        ops[parts[0].split(".")[0]] += int(r[iE])
        tot += int(r[iE])
    print("SASS opcode mix (warp instructions executed, first kernel in the report):\n")
    print("| opcode | executed | share |\n|---|---|---|")
    for op, n in ops.most_common(16):
        print("| %s | %d | %.1f %% |" % (op, n, 100.0 * n / tot))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        full(sys.argv[2])
