#!/usr/bin/env python
"""Per-region executed instructions and stall samples of a kernel from an .ncu-rep (SASS page),
regions = maximal runs between branch targets, merged into the loops they belong to; prints
instructions per update.   usage: tools/ncu_regions.py rep n_updates [bin]"""
import csv, io, re, subprocess, sys
import os
rep, nupd = os.path.abspath(sys.argv[1]), float(sys.argv[2])
binsz = int(sys.argv[3]) if len(sys.argv) > 3 else 64
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, cwd="/tmp").stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
base = int(body[0][ix["Address"]], 16)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[ix["# Samples"]]) for r in body)
print("total: %.1f instr/update, %d samples" % (tot_i / nupd, tot_s))
# bins of `binsz` instructions
for b0 in range(0, len(body), binsz):
    chunk = body[b0:b0 + binsz]
    ie = sum(int(r[ix["Instructions Executed"]]) for r in chunk)
    sm = sum(int(r[ix["# Samples"]]) for r in chunk)
    if ie / nupd < 2 and sm / tot_s < 0.005:
        continue
    ops = {}
    for r in chunk:
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    a0 = int(chunk[0][ix["Address"]], 16) - base
    print("%#7x +%3d: %7.1f instr/upd  %5.1f%% samples   %s" % (
        a0, len(chunk), ie / nupd, 100.0 * sm / tot_s, " ".join("%s:%.0f" % (k, v / nupd) for k, v in top)))
