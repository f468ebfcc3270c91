"""The C host library under the sanitizers, driven by a plain C caller (tests/c/host_api_check.c).

SURVEY.md section 5 lists sanitizers among the reference's auxiliary practices; compute-sanitizer is not usable on the GPU
pool, so the device side is covered by the CPU emulator (tests/test_emulator.py) and THIS covers the host side: the
library's own C sources (instance init, dispatcher, staging pipeline, plan caches, thread-local pools) compiled with
-fsanitize=address,undefined and with -fsanitize=thread, linked against the real CUDA shim.

  not gpu: no device -> every batched call must refuse with ARM_MATH_CUDA_NO_DEVICE and write nothing;
  gpu:     one worker / three workers on one device / four host threads / small staging chunks -> identical bits.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("host_c_build", os.path.join(HERE, "c", "build.py"))   # tests/emu has a build.py too
cbuild = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(cbuild)

# CUDA maps its own address ranges: ASan must not protect the shadow gap; the driver's process-lifetime allocations are
# not this library's leaks (the library's own are still reported: its frees happen before exit)
ENV = {"asan": {"ASAN_OPTIONS": "protect_shadow_gap=0:detect_leaks=0:abort_on_error=0", "UBSAN_OPTIONS": "print_stacktrace=1:halt_on_error=1"},
       "tsan": {"TSAN_OPTIONS": "halt_on_error=0:report_signal_unsafe=0:exitcode=66"}}


def binaries():
    have_build_inputs = os.path.exists(os.path.join(ROOT, "cmsis-dsp_b200", "build", "cmsisdsp_tables_generated.c"))
    if have_build_inputs:
        cbuild.build()
    out = {v: cbuild.binary(v) for v in cbuild.VARIANTS}
    for v, p in out.items():
        assert os.path.exists(p), f"{p} is missing: run python -c 'import __graft_entry__ as g; g.build()'"
    return out


def run(variant, env_extra=None):
    env = dict(os.environ, **ENV[variant], **(env_extra or {}))
    r = subprocess.run([binaries()[variant]], capture_output=True, text=True, env=env, timeout=600)
    return r


@pytest.mark.parametrize("variant", ["asan", "tsan"])
def test_no_device_every_call_refuses_cleanly(variant):
    r = run(variant, {"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr and "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]


@pytest.mark.gpu
def test_dispatcher_under_asan_ubsan_on_the_device():
    r = run("asan")
    assert r.returncode == 0, r.stdout + r.stderr[-4000:]
    assert "same bits" in r.stdout, r.stdout
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]


@pytest.mark.gpu
def test_dispatcher_under_tsan_on_the_device():
    r = run("tsan")
    assert "same bits" in r.stdout, r.stdout + r.stderr[-4000:]
    own = [blk for blk in r.stderr.split("==================") if "WARNING: ThreadSanitizer" in blk and race_is_in_own_code(blk)]
    assert not own, own[0][-3000:]


OWN_FILES = ("arm_cuda_engine.c", "arm_cfft_exec.c", "arm_mfcc.c", "arm_cfft_init.c", "arm_cfft_deprecated.c", "host_api_check.c")


def race_is_in_own_code(block):
    """A report counts when one of the two racing ACCESSES is made by the library's own code: the innermost frame of an
    access stack, below the sanitizer's interceptors, is one of our sources.  Accesses made inside the uninstrumented CUDA
    driver (e.g. its memcpy into a command buffer of its own, reached through cudaMemcpyAsync from two workers) are the
    driver's business: it synchronises them in ways TSan cannot see, and our frames merely appear further up the stack."""
    for section in block.split("\n\n"):
        head = section.lstrip().split("\n", 1)[0]
        if not any(head.startswith(k) for k in ("WARNING", "Write of", "Read of", "Previous write", "Previous read", "Atomic", "Previous atomic")):
            continue
        frames = [ln for ln in section.splitlines() if ln.strip().startswith("#")]
        frames = [ln for ln in frames if "libtsan" not in ln and "sanitizer_common" not in ln]
        if frames and any(f in frames[0] for f in OWN_FILES):
            return True
    return False


def test_tsan_report_filter():
    theirs = """WARNING: ThreadSanitizer: data race (pid=1)
  Write of size 8 at 0x1 by main thread:
    #0 memcpy ../sanitizer_common/x.inc:115 (libtsan.so.2+0x8bd30)
    #1 <null> <null> (libcuda.so.1+0x28efa4)
    #2 worker /root/repo/cmsis-dsp_b200/csrc/host/arm_cuda_engine.c:294 (host_api_check_tsan+0xb142)

  Previous write of size 8 at 0x1 by thread T7:
    #0 memcpy ../sanitizer_common/x.inc:115 (libtsan.so.2+0x8bd30)
    #1 <null> <null> (libcuda.so.1+0x28efa4)
    #2 worker /root/repo/cmsis-dsp_b200/csrc/host/arm_cuda_engine.c:294 (host_api_check_tsan+0xb142)
"""
    ours = theirs.replace("    #1 <null> <null> (libcuda.so.1+0x28efa4)\n", "", 1).replace("#0 memcpy ../sanitizer_common/x.inc:115 (libtsan.so.2+0x8bd30)\n    #2 worker", "#0 worker", 1)
    assert not race_is_in_own_code(theirs)
    assert race_is_in_own_code(ours)
