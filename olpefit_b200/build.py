"""Build liblapf.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "liblapf.so")
SOURCES = [os.path.join(CSRC, "lapf.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "lapf_device.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "lapf.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; liblapf.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # build next to the target and rename: several ranks may decide to rebuild at the same time
    tmp = "%s.tmp.%d" % (LIB, os.getpid())
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-shared", "-Xcompiler", "-fPIC", "-o", tmp] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    # which toolchain produced this binary: the sampler kernel is sensitive to it, the
    # self-test at sampler creation decides whether the build is usable, this file says what was used
    try:
        ver = subprocess.run([cmd[0], "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2:]
        with open(LIB + ".buildinfo", "w") as fh:
            fh.write("nvcc: %s\nflags: %s\n" % (" | ".join(ver), " ".join(cmd[1:-len(SOURCES) - 2])))
    except Exception:
        pass
    return LIB


def build_info() -> str:
    """nvcc version and flags of the library on disk ('' if it was not built by this module)."""
    try:
        with open(LIB + ".buildinfo") as fh:
            return fh.read()
    except OSError:
        return ""


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
