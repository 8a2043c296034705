#!/usr/bin/env python
"""List every innermost loop (backward branch) of a kernel with its instruction mix.
usage: tools/sass_loops.py <mangled-name-substring> [min_instructions]"""
import re, subprocess, sys, collections
lib = "olpefit_b200/csrc/liblapf.so"
pat = sys.argv[1]
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = [b for b in blocks if pat in b.split("\n")[0]][0]
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:!?U?P[T0-9]+,\s*)?`?\(?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] <= i:
            loops.append((addr[tgt], i))
# innermost only
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
print("kernel instructions: %d, loops: %d (innermost %d)" % (len(ins), len(loops), len(inner)))
for lo, hi in inner:
    n = hi - lo + 1
    if n < minlen:
        continue
    ops = collections.Counter()
    for _, t in ins[lo:hi + 1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0]] += 1
    packed = sum(v for k, v in ops.items() if k.startswith(("FFMA2", "FMUL2", "FADD2")))
    scalar = sum(v for k, v in ops.items() if k.split(".")[0] in ("FFMA", "FMUL", "FADD"))
    print("loop %#x..%#x: %d instr; packed %d scalar-fp %d MUFU %d -> FMA-pipe cycles %d" % (
        ins[lo][0], ins[hi][0], n, packed, scalar, ops["MUFU.EX2"], 2 * packed + scalar))
    print("   " + "  ".join("%s:%d" % kv for kv in ops.most_common()))
