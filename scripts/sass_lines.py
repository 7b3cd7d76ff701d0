"""Attribute ncu per-SASS-instruction counters to CUDA source lines (runs without a GPU).

usage: python scripts/sass_lines.py <ncu source-page csv> <nvdisasm -g -c listing> <kernel substring> [top]
  ncu -i prof.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all obj.o; nvdisasm -g -c obj.sm_100a.cubin > dis.txt
"""
import collections
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith(".text.")]
sel = [i for i in starts if kern in lines[i]]
assert len(sel) == 1, "kernel substring must match exactly one .text section: %d" % len(sel)
beg = sel[0]
end = min([i for i in starts if i > beg] + [len(lines)])
cur, inst = None, []
for l in lines[beg:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        inst.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
assert len(data) == len(inst), (len(data), len(inst))
FP = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
by, fp, smp, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for (key, text), r in zip(inst, data):
    parts = text.split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    n = int(r[iE])
    by[key] += n
    smp[key] += int(r[iN])
    ops[key][op] += n
    if op in FP:
        fp[key] += n
tot, ts = sum(by.values()), sum(smp.values())
print("total warp instructions %d, samples %d" % (tot, ts))
for key, v in by.most_common(top):
    f, l = key if key else ("?", 0)
    mix = " ".join("%s:%.1f" % (o, 100.0 * c / tot) for o, c in ops[key].most_common(4))
    print("%-18s:%4d inst %5.1f%% fp64 %4.1f%% stall-samples %5.1f%%  %s" % (f, l, 100.0 * v / tot, 100.0 * fp[key] / tot,
                                                                      100.0 * smp[key] / ts, mix))
