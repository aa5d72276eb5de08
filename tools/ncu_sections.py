#!/usr/bin/env python3
"""Per-section instruction histogram of a kernel from tools/ncu_sass_hot.py's full listing: SASS
instructions grouped by how often they ran (each distinct execution count is one nesting level of the
kernel -- per warp, per work item, per source row, per inner-loop trip), with the share of all warp
instructions each level accounts for and its opcode mix."""
import collections, sys
rows = []
total = 0
for ln in open(sys.argv[1]):
    f = ln.split()
    if ln.startswith("total warp instructions"):
        total = int(f[-1])
        continue
    if len(f) < 5 or not f[0].isdigit():
        continue
    n = int(f[0])
    ops = [t for t in f[4:] if not t.startswith("@")]
    rows.append((n, ops[0].split(".")[0] if ops else "?"))
levels = collections.defaultdict(list)
for n, op in rows:
    if n:
        levels[n].append(op)
print("total warp instructions", total)
print("%12s %7s %14s %7s  %s" % ("executions", "instrs", "warp instrs", "share", "opcode mix"))
for n in sorted(levels, reverse=True):
    ops = levels[n]
    mix = collections.Counter(ops).most_common(6)
    w = n * len(ops)
    print("%12d %7d %14d %6.1f%%  %s" % (n, len(ops), w, 100.0 * w / max(total, 1), " ".join("%s:%d" % m for m in mix)))
