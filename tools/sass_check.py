#!/usr/bin/env python
"""Static checks of the shipped SASS (cuobjdump) that the tests run on every build.

1. TMEM loads.  `tcgen05.ld` (SASS LDTM) is issued by one asm statement and waited for by another
   (`tcgen05.wait::ld`, csrc/lapf_device.cuh); ptxas turns the wait into a scoreboard dependency.
   For every LDTM the checker decodes its write-barrier slot from the control bits and walks
   forward: the first instruction that touches one of the destination registers must wait on
   that slot (directly, or after an earlier instruction already waited on it).
2. Convergence.  LDTM is `.sync.aligned`: the whole warp has to execute it together.  Every LDTM of
   the sampler kernels must be reached through a WARPSYNC.ALL / BRA.DIV that is not followed by a
   potentially divergent region, or sit in a loop whose body re-converges the warp (WARPSYNC.ALL
   at its top).  The builds that failed in round 1 / 2 had neither.

Control word of an sm_70+ instruction (bits of the 128-bit encoding): stall 105-108, yield 109,
write barrier 110-112, read barrier 113-115, wait mask 116-121, reuse 122-125.

    python tools/sass_check.py [lib.so]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "olpefit_b200", "csrc", "liblapf.so")


class Ins:
    __slots__ = ("addr", "text", "lo", "hi")

    def __init__(self, addr, text, lo, hi):
        self.addr, self.text, self.lo, self.hi = addr, text, lo, hi

    @property
    def op(self):
        return re.sub(r"^@!?U?P\d+\s+", "", self.text).split()[0]

    @property
    def stall(self):
        return (self.hi >> 41) & 0xF

    @property
    def wbar(self):
        return (self.hi >> 46) & 7

    @property
    def rbar(self):
        return (self.hi >> 49) & 7

    @property
    def wait(self):
        return (self.hi >> 52) & 0x3F

    def regs(self):
        """Vector register numbers named by the instruction, with the widths the mnemonic implies."""
        out = set()
        txt = re.sub(r"^@!?U?P\d+\s+", "", self.text)
        width = 1
        m = re.search(r"\.(64|128)\b", txt.split()[0])
        if m:
            width = int(m.group(1)) // 32
        for m in re.finditer(r"(?<![U\w])R(\d+)(\.F32x2\.HI_LO|\.64)?", txt):
            n = int(m.group(1))
            w = 2 if m.group(2) else 1
            for k in range(max(w, 1)):
                out.add(n + k)
        if width > 1:      # vector loads / stores: the DATA operand spans `width` registers, addresses do not
            names = [int(m.group(1)) for m in re.finditer(r"(?<![U\w])R(\d+)", txt)]
            if names:
                data = names[0] if self.op.startswith("LD") else names[-1]
                for k in range(width):
                    out.add(data + k)
        if self.op.startswith("DADD") or self.op.startswith("DMUL") or self.op.startswith("DFMA") or "F64" in self.op:
            for m in re.finditer(r"(?<![U\w])R(\d+)", txt):
                out.add(int(m.group(1)) + 1)
        return out


def functions(lib=DEFAULT_LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = {}
    for blk in out.split("Function : ")[1:]:
        name = blk.split("\n")[0].strip()
        ins, lines = [], blk.split("\n")
        for i, ln in enumerate(lines):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", ln)
            if m:
                m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1]) if i + 1 < len(lines) else None
                ins.append(Ins(int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), int(m2.group(1), 16) if m2 else 0))
        res[name] = ins
    return res


def check_ldtm_waits(ins, horizon=400):
    """[(address, message)] for LDTM whose destination is touched before its scoreboard slot is waited for."""
    bad, gaps = [], []
    for i, x in enumerate(ins):
        if not x.op.startswith("LDTM"):
            continue
        m = re.match(r"LDTM\.x(\d+)\s+R(\d+)", re.sub(r"^@!?U?P\d+\s+", "", x.text))
        n, first = int(m.group(1)), int(m.group(2))
        dest = set(range(first, first + n))
        slot = x.wbar
        if slot == 7:
            bad.append((x.addr, "LDTM without a write barrier"))
            continue
        for j in range(i + 1, min(len(ins), i + 1 + horizon)):
            y = ins[j]
            if y.wait & (1 << slot):
                gaps.append(j - i)
                break
            if y.regs() & dest:
                bad.append((x.addr, "%#x %s touches R%d..R%d before scoreboard %d is waited for" % (
                    y.addr, y.text, first, first + n - 1, slot)))
                break
            if y.op in ("EXIT", "RET"):
                break
    return bad, gaps


def check_ldtm_convergence(ins):
    """[(address, message)] for LDTM not covered by an unconditional convergence point.

    Walking backwards from the LDTM through straight-line code: a WARPSYNC.ALL (or a CTA barrier, or an
    earlier LDTM of the same straight line) must come BEFORE the nearest BRA.DIV.  BRA.DIV is how
    ptxas lowers __syncwarp() when it expects the warp to be converged: one run-time check, after
    which every later __syncwarp() of the region becomes a NOP.  The builds whose sampler evaluated
    the update after a recorded update wrongly reached all their TMEM loads through
    such a check alone; every build that works re-converges the warp with WARPSYNC.ALL in between
    (at the top of each pixel loop, or before the first load of an unrolled one)."""
    bad = []
    for i, x in enumerate(ins):
        if not x.op.startswith("LDTM"):
            continue
        ok, why = False, "start of the kernel"
        j = i - 1
        while j >= 0:
            y = ins[j]
            if y.op.startswith("WARPSYNC.ALL") or y.op.startswith("BAR") or y.op.startswith("LDTM"):
                ok = True
                break
            if y.op.startswith("BRA.DIV"):
                why = "BRA.DIV at %#x" % y.addr
                break
            if y.op in ("EXIT",) or (y.op == "BRA" and not y.text.startswith("@")):
                why = "unconditional jump at %#x" % y.addr
                break
            j -= 1
        if not ok:
            bad.append((x.addr, "no WARPSYNC.ALL between this LDTM and the %s" % why))
    return bad


def summary(name, ins):
    ops = {}
    for x in ins:
        ops[x.op] = ops.get(x.op, 0) + 1
    return {k: ops.get(k, 0) for k in ("LDTM.x16", "LDTM.x8", "STTM.x16", "STTM.x8", "UBLKCP.S.G", "WARPSYNC.ALL", "BRA.DIV",
                                       "FFMA2", "FMUL2", "MUFU.EX2", "NOP")}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_LIB
    rc = 0
    for name, ins in sorted(functions(lib).items()):
        if "gibbs_batch_kernel" not in name:
            continue
        bad_w, gaps = check_ldtm_waits(ins)
        bad_c = check_ldtm_convergence(ins)
        s = summary(name, ins)
        print("%s: %d instructions, %s" % (name, len(ins), " ".join("%s:%d" % kv for kv in s.items() if kv[1])))
        print("   LDTM -> first wait distance (instructions): min %d median %d max %d" % (
            min(gaps), sorted(gaps)[len(gaps) // 2], max(gaps)) if gaps else "   no LDTM")
        for a, msg in bad_w + bad_c:
            print("   VIOLATION at %#x: %s" % (a, msg))
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
