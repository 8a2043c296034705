#!/usr/bin/env python
"""Static instruction count of one kernel by source line (nvdisasm -g), grouped into the phases of an
update -- where the instructions of a batched-sampler round sit outside the pixel loop.
usage: tools/sass_lines.py <lib.so> <mangled-name-substring> [--lines]"""
import collections, os, re, subprocess, sys, tempfile


def kernel_lines(lib, pattern):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True, check=True).stdout
    sect = None
    cur = ("?", 0)
    rows = []
    for ln in out.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect = m.group(1)
            continue
        if sect is None or pattern not in sect:
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)( inlined at "(.*?)", line (\d+))?', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return rows


if __name__ == "__main__":
    rows = kernel_lines(sys.argv[1], sys.argv[2])
    by = collections.Counter()
    mufu = collections.Counter()
    for a, t, (f, l) in rows:
        by[(f, l)] += 1
        if "MUFU" in t:
            mufu[(f, l)] += 1
    print("%d instructions" % len(rows))
    if "--lines" in sys.argv:
        for (f, l), n in sorted(by.items()):
            print("%-22s %5d  %4d instr  %3d MUFU" % (f, l, n, mufu[(f, l)]))
