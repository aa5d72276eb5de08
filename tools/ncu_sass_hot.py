#!/usr/bin/env python3
"""Per-SASS-instruction executed counts from an .ncu-rep (source page), grouped into address ranges."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iex]) for r in rows[2:] if len(r) > iex and r[iex].isdigit())
print('total warp instructions', tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    v = int(r[iex])
    if top == 0 or v * 400 > tot:
        print('%8d %5.1f%% %6s  %s  %s' % (v, 100.0 * v / tot, r[ist], r[ia][-5:], r[isrc][:90]))
