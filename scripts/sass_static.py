"""Static SASS count of a kernel's main loop by source line / opcode (no GPU, no ncu).
usage: python scripts/sass_static.py <nvdisasm -g -c listing> <kernel substring>
Counts the instructions inside the address range of the widest backward branch (the plane march)."""
import collections, re, sys
dis, kern = sys.argv[1], sys.argv[2]
lines = open(dis).read().split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith(".text.")]
sel = [i for i in starts if kern in lines[i]]
assert len(sel) == 1, len(sel)
beg = sel[0]; end = min([i for i in starts if i > beg] + [len(lines)])
cur = None; inst = []; labels = {}
for l in lines[beg:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        labels[m.group(1)] = len(inst); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        inst.append((cur, m.group(2)))
best = None
for n, (k, t) in enumerate(inst):
    m = re.search(r"BRA.*?(\.L_x_\d+)", t)
    if m and m.group(1) in labels and labels[m.group(1)] < n:
        span = (labels[m.group(1)], n)
        if best is None or span[1] - span[0] > best[1] - best[0]:
            best = span
lo, hi = best
body = inst[lo:hi + 1]
print("loop body: %d instructions (of %d in kernel)" % (len(body), len(inst)))
FP = ("DADD", "DMUL", "DFMA", "DSETP")
byline, fpl, ops = collections.Counter(), collections.Counter(), collections.Counter()
for k, t in body:
    parts = t.split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    ops[op] += 1
    byline[k] += 1
    if op in FP: fpl[k] += 1
print("opcode mix:", " ".join("%s:%d" % kv for kv in ops.most_common(24)))
print("fp64 total:", sum(ops[o] for o in FP))
for k, v in sorted(byline.items(), key=lambda kv: (kv[0] or ("?", 0))):
    print("%-20s:%4d  inst %4d  fp64 %4d" % ((k or ("?", 0))[0], (k or ("?", 0))[1], v, fpl[k]))
