"""Builds the CPU emulation of the CUDA sources (TEST INFRASTRUCTURE ONLY; see cs_emu.h).

g++ compiles co_snarks_b200/csrc/*.cu as C++ with -DCS_EMU so the exact device algorithms can be
checked against the oracle on a box without a GPU.  The product never loads this library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "co_snarks_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libcosnarks_emu.so")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")] + [os.path.join(HERE, "cs_emu.cpp")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cs_emu.h", "cs_emu.cpp")]
    deps.append(os.path.join(ROOT, "include", "cosnarks_gpu.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g0", "-DCS_EMU", "-DCS_ENABLE_BLS12_381", "-fPIC", "-shared", "-pthread", "-w",
           "-I", HERE, "-I", CSRC, "-o", OUT]
    for s in sources():
        cmd += ["-x", "c++", s]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
