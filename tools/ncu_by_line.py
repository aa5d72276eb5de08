#!/usr/bin/env python3
"""Executed warp instructions of one kernel attributed to SOURCE LINES: joins the per-instruction counts
of tools/ncu_sass_hot.py's full listing (profiles/rNN_ncu_sass_all_<cfg>.txt) with the line table of the
built library (cuobjdump -xelf + nvdisasm -g; the kernel TU is compiled with -lineinfo).  The library must
be the build the profile was taken from.
Usage: ncu_by_line.py <sass_all listing> <mangled kernel name> [top N]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
listing, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "smolscale_b200", "libsmolscale_cuda.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if "kernels" in f and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, ln in enumerate(sass) if ln.startswith("//---------------------") and ".text." + kernel in ln][0]
cur, ins = None, []
for ln in sass[start + 1:]:
    if ln.startswith("//---------------------") and ins:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+.*?;", ln):
        ins.append(cur)
cnt = [int(f[0]) for f in (ln.split() for ln in open(listing)) if len(f) >= 5 and f[0].isdigit()]
if len(cnt) != len(ins):
    sys.exit("instruction counts differ (%d in the profile, %d in the library): not the same build" % (len(cnt), len(ins)))
by = collections.Counter()
for line, c in zip(ins, cnt):
    by[line] += c
tot = sum(by.values())
src = open(os.path.join(ROOT, "smolscale_b200", "csrc", "smolscale-cuda-kernels.cu")).read().split("\n")
print("total warp instructions", tot)
for line, c in by.most_common(top):
    print("%6.2f%% %10d  L%-5s %s" % (100.0 * c / tot, c, line, src[line - 1].strip()[:110] if line else ""))
