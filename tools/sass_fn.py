#!/usr/bin/env python
"""Print the SASS of one kernel of a library: tools/sass_fn.py <lib.so> <mangled-name-substring>"""
import re
import subprocess
import sys


def function_sass(lib, pattern):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    for blk in out.split("Function : ")[1:]:
        if pattern in blk.split("\n")[0]:
            ins = []
            for ln in blk.split("\n"):
                m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip()))
            return ins
    raise KeyError(pattern)


if __name__ == "__main__":
    for a, t in function_sass(sys.argv[1], sys.argv[2]):
        print("%05x  %s" % (a, t))
