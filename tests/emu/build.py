#!/usr/bin/env python
"""Builds tests/emu/libemu.so (the CPU emulator of the kernel bodies; TEST INFRASTRUCTURE) in parallel parts.
Used by __graft_entry__.build() and by tests/test_emulator.py when a kernel header is newer than the library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "..", "cmsis-dsp_b200", "csrc", "cuda"))
PARTS = 5


def deps():
    return [os.path.join(HERE, "emu.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]


def stale():
    so = os.path.join(HERE, "libemu.so")
    return not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps())


def build(force=False):
    so = os.path.join(HERE, "libemu.so")
    if not force and not stale():
        return so
    objs = [os.path.join(HERE, f"emu_part{k}.o") for k in range(PARTS)]

    def one(k):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-Wno-unknown-pragmas", f"-DEMU_PART={k}", "-I", CSRC,
                               "-c", os.path.join(HERE, "emu.cpp"), "-o", objs[k]])
    with ThreadPoolExecutor(max_workers=PARTS) as ex:
        list(ex.map(one, range(PARTS)))
    subprocess.check_call(["g++", "-shared", "-o", so] + objs)
    for o in objs:
        os.remove(o)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
