#!/usr/bin/env python
"""Stall samples and executed instructions of a kernel (ncu --set full report) by source line:
SASS addresses from the report, lines from nvdisasm -g of the library the report was taken with.
usage: tools/ncu_lines.py rep lib.so kernel-substring [top]"""
import collections, csv, io, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import kernel_lines

rep, lib, pat = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, cwd="/tmp").stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
base = int(body[0][ix["Address"]], 16)
line_of = {a: (t, fl) for a, t, fl in kernel_lines(lib, pat)}
funcs = []
src = os.path.join(os.path.dirname(os.path.abspath(lib)), "lapf_device.cuh")
for i, l in enumerate(open(src), 1):
    m = re.match(r"__device__ __forceinline__ \S+.*? (\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))

def where(f, l):
    if f == "lapf_device.cuh":
        name = "?"
        for i, n in funcs:
            if i <= l:
                name = n
        return name
    return "%s:%d" % (f, l)

samp, inst = collections.Counter(), collections.Counter()
tot = 0
for r in body:
    a = int(r[ix["Address"]], 16) - base
    t, (f, l) = line_of.get(a, ("?", ("?", 0)))
    k = where(f, l)
    s = int(r[ix["# Samples"]])
    samp[k] += s
    inst[k] += int(r[ix["Instructions Executed"]])
    tot += s
print("%d samples" % tot)
for k, s in samp.most_common(top):
    print("%-34s %6.2f%%  %10d instr" % (k, 100.0 * s / tot, inst[k]))
