#!/usr/bin/env python
"""Builds tests/c/host_api_check.c together with the library's own host sources under the sanitizers:
    cmsis-dsp_b200/lib/host_api_check_asan   -fsanitize=address,undefined
    cmsis-dsp_b200/lib/host_api_check_tsan   -fsanitize=thread
The binaries sit next to libcmsisdsp_cuda.so (rpath $ORIGIN) so that they travel to the GPU box with the built libraries
(cmsis-dsp_b200/build/, where the generated table source lives, does not).  Needs `make -C cmsis-dsp_b200/csrc` first."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "cmsis-dsp_b200")
LIB = os.path.join(PKG, "lib")
VARIANTS = {"asan": ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], "tsan": ["-fsanitize=thread"]}


def sources():
    host = sorted(glob.glob(os.path.join(PKG, "csrc", "host", "*.c")))
    return host + [os.path.join(PKG, "build", "cmsisdsp_tables_generated.c"), os.path.join(ROOT, "tests", "c", "host_api_check.c")]


def binary(variant):
    return os.path.join(LIB, f"host_api_check_{variant}")


def build(force=False):
    srcs = sources()
    for v, flags in VARIANTS.items():
        out = binary(v)
        if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
            continue
        missing = [s for s in srcs if not os.path.exists(s)]
        if missing:
            raise RuntimeError("run `make -C cmsis-dsp_b200/csrc all` first: missing " + ", ".join(missing))
        cmd = ["gcc", "-std=gnu11", "-O1", "-g", "-fno-omit-frame-pointer", "-Wall", "-Wextra", *flags,
               f"-I{ROOT}/include", f"-I{PKG}/csrc/cuda", f"-I{PKG}/csrc/host", *srcs, "-o", out,
               f"-L{LIB}", "-lcmsisdsp_cuda", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"]
        subprocess.check_call(cmd)
    return [binary(v) for v in VARIANTS]


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv)))
