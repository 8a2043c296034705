#!/usr/bin/env python
"""Dispatch-cost estimate of the pixel loops of a sampler kernel from its SASS, with the per-instruction
costs measured by tools/microbench7.cu on B200 (profiles/r02_microbench7.txt):

  packed FP32 (FFMA2 / FMUL2 / FADD2)  max(2, operands that must be fetched from the register file) + 0.15
  scalar FP32 (FFMA / FMUL / FADD)     1.0 + 0.32 per fetched operand beyond the first
  MUFU                                 1.0      LDS / LDTM / STS   1.0
  integer / logic / moves (ALU pipe)   1.4      everything else    1.0

An operand is "fetched" unless the previous instruction kept the same register in the same operand
slot with the .reuse flag (operand-reuse cache).  Usage: tools/sass_cost.py [lib.so] [kernel-substring]
Prints every innermost loop that contains a TMEM load: instructions, FP32-pipe cycles (2 per packed,
1 per scalar), estimated dispatch cycles, per trip."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_check  # noqa: E402

PACKED = ("FFMA2", "FMUL2", "FADD2")
SCALAR = ("FFMA", "FMUL", "FADD")
ALU = ("IADD3", "LOP3", "SHF", "ISETP", "FSEL", "SEL", "MOV", "IMAD.MOV", "VIADD", "LEA", "I2FP", "PRMT", "R2UR", "UMOV",
       "IMAD", "FSETP", "R2P", "PLOP3", "CS2R", "S2R")


def operands(text):
    txt = re.sub(r"^@!?U?P\d+\s+", "", text)
    op = txt.split()[0]
    rest = txt[len(op):]
    return op, [o.strip() for o in rest.split(",")]


def cost(prev, ins):
    op, ops = operands(ins.text)
    srcs = ops[1:]
    pops = operands(prev.text)[1][1:] if prev is not None else []
    need = set()
    for slot, o in enumerate(srcs):
        m = re.match(r"-?\|?(R\d+)", o)
        if not m or m.group(1) == "RZ":
            continue                                     # immediates, uniform registers, constants: no vector fetch
        reused = slot < len(pops) and ".reuse" in pops[slot] and re.match(r"-?\|?(R\d+)", pops[slot]) and \
            re.match(r"-?\|?(R\d+)", pops[slot]).group(1) == m.group(1)
        if not reused:
            need.add(m.group(1))                         # the same register in two slots is fetched once
    fetched = len(need)
    if op.startswith(PACKED):
        return max(2.0, float(fetched)) + 0.15, 2.0
    if op.startswith(SCALAR):
        return 1.0 + 0.32 * max(0, fetched - 1), 1.0
    if op.startswith("MUFU") or op.startswith(("LDS", "LDTM", "STS", "LDG", "STG")):
        return 1.0, 0.0
    if op.startswith(ALU):
        return 1.4, 0.0
    return 1.0, 0.0


def loops_with_ldtm(ins):
    addr = {x.addr: i for i, x in enumerate(ins)}
    out = []
    for i, x in enumerate(ins):
        m = re.search(r"BRA\s+(?:!?U?P[0-9T]+,\s*)?0x([0-9a-f]+)", x.text)
        if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
            lo = addr[int(m.group(1), 16)]
            body = ins[lo:i + 1]
            inner = not any(re.search(r"BRA\s+(?:!?U?P[0-9T]+,\s*)?0x([0-9a-f]+)", y.text) and
                            int(re.search(r"0x([0-9a-f]+)", y.text).group(1), 16) in addr and
                            lo < addr[int(re.search(r"0x([0-9a-f]+)", y.text).group(1), 16)] < addr[y.addr]
                            for y in body[:-1])
            if inner and any(y.op.startswith("LDTM") for y in body):
                out.append((lo, i))
    return out


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else sass_check.DEFAULT_LIB
    pat = sys.argv[2] if len(sys.argv) > 2 else "gibbs_batch_kernelILi2ELi64ELi64"
    for name, ins in sass_check.functions(lib).items():
        if pat not in name:
            continue
        print(name)
        for lo, hi in loops_with_ldtm(ins):
            body = ins[lo:hi + 1]
            steps = sum(y.op.startswith("LDTM") for y in body)
            tot = pipe = 0.0
            prev = ins[lo - 1] if lo else None
            for y in body:
                c, p = cost(prev, y)
                tot += c
                pipe += p
                prev = y
            mufu = sum(y.op.startswith("MUFU") for y in body)
            packed = sum(y.op.startswith(PACKED) for y in body)
            print("  loop %#x..%#x: %3d instructions for %d warp step(s): %2d packed, %d MUFU; FP32 pipe %.0f cycles/step, "
                  "estimated dispatch %.0f cycles/step" % (ins[lo].addr, ins[hi].addr, len(body), steps, packed // steps,
                                                            mufu // steps, pipe / steps, tot / steps))


if __name__ == "__main__":
    main()
