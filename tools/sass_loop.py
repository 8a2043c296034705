#!/usr/bin/env python
"""Print the instruction mix of the pixel loop (the innermost loop containing MUFU.EX2) of a kernel.
usage: tools/sass_loop.py <mangled-name-substring> [--list]"""
import re, subprocess, sys, collections
lib = "olpefit_b200/csrc/liblapf.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = [b for b in blocks if pat in b.split("\n")[0]][0]
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
ex = [i for i, (_, t) in enumerate(ins) if "MUFU.EX2" in t]
# innermost backward branch enclosing the first EX2
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:U[P0-9!]+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] <= ex[0] <= i:
            if best is None or (i - addr[tgt]) < (best[1] - best[0]):
                best = (addr[tgt], i)
lo, hi = best
ops = collections.Counter()
for _, t in ins[lo:hi + 1]:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t.split()[0]] += 1
n = hi - lo + 1
nex = ops["MUFU.EX2"]
print("loop %#x..%#x: %d instructions, %d MUFU.EX2 -> %.2f issue slots per EX2 (budget 8)" % (ins[lo][0], ins[hi][0], n, nex, n / nex))
print("  " + "  ".join("%s:%d" % kv for kv in ops.most_common()))
print("  total EX2 in kernel: %d, kernel instructions: %d" % (len(ex), len(ins)))
if "--list" in sys.argv:
    for a, t in ins[lo:hi + 1]:
        print("%05x  %s" % (a, t))
