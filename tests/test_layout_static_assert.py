"""Compile-time layout check of include/*.h against the reference's headers.

The reference's instance structs are the ABI of the drop-in boundary: a caller compiled against the reference's
`dsp/transform_functions.h` hands these structs to this library.  Two steps:

1. where /root/reference exists (the build container), a probe TU compiled against the REFERENCE headers
   (generic branch: -D__GNUC_PYTHON__, as the oracle build) prints sizeof / offsetof of every instance struct and the
   arm_status values; the result must equal the committed tests/golden/ref_layout.json (made by this very probe:
   `python tests/test_layout_static_assert.py --write`);
2. everywhere, a TU that includes the REPO's headers is compiled with one _Static_assert per number of that file.
   A layout drift is a compile error.
"""
import json
import os
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLDEN = os.path.join(HERE, "golden", "ref_layout.json")

CFFT = ["fftLen", "pTwiddle", "pBitRevTable", "bitRevLength"]
RADIX = ["fftLen", "ifftFlag", "bitReverseFlag", "pTwiddle", "pBitRevTable", "twidCoefModifier", "bitRevFactor"]
RFIX = ["fftLenReal", "ifftFlagR", "bitReverseFlagR", "twidCoefRModifier", "pTwiddleAReal", "pTwiddleBReal", "pCfft"]
STRUCTS = {
    "arm_cfft_instance_q15": CFFT, "arm_cfft_instance_q31": CFFT, "arm_cfft_instance_f32": CFFT, "arm_cfft_instance_f64": CFFT,
    "arm_rfft_fast_instance_f32": ["Sint", "fftLenRFFT", "pTwiddleRFFT"],
    "arm_rfft_fast_instance_f64": ["Sint", "fftLenRFFT", "pTwiddleRFFT"],
    "arm_rfft_instance_q15": RFIX, "arm_rfft_instance_q31": RFIX,
    "arm_cfft_radix4_instance_q15": RADIX, "arm_cfft_radix4_instance_q31": RADIX,
    "arm_cfft_radix4_instance_f32": RADIX + ["onebyfftLen"], "arm_cfft_radix2_instance_f32": RADIX + ["onebyfftLen"],
    "arm_cfft_radix2_instance_q15": RADIX, "arm_cfft_radix2_instance_q31": RADIX,
    "arm_mfcc_instance_f32": ["dctCoefs", "filterCoefs", "windowCoefs", "filterPos", "filterLengths", "fftLen", "nbMelFilters",
                              "nbDctOutputs", "rfft"],
}
STATUS = ["ARM_MATH_SUCCESS", "ARM_MATH_ARGUMENT_ERROR", "ARM_MATH_LENGTH_ERROR", "ARM_MATH_SIZE_MISMATCH", "ARM_MATH_NANINF",
          "ARM_MATH_SINGULAR", "ARM_MATH_TEST_FAILURE", "ARM_MATH_DECOMPOSITION_FAILURE"]
SCALARS = ["q15_t", "q31_t", "float32_t", "float64_t", "arm_status"]


def probe_reference():
    """sizeof / offsetof / enum values as the REFERENCE's headers define them"""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "arm_math_types.h"', '#include "dsp/transform_functions.h"',
             'int main(void) {']
    for s, fields in STRUCTS.items():
        lines.append(f'  printf("sizeof/{s} %zu\\n", sizeof({s}));')
        for f in fields:
            lines.append(f'  printf("offsetof/{s}/{f} %zu\\n", offsetof({s}, {f}));')
    for t in SCALARS:
        lines.append(f'  printf("sizeof/{t} %zu\\n", sizeof({t}));')
    for e in STATUS:
        lines.append(f'  printf("enum/{e} %d\\n", (int){e});')
    lines.append("  return 0; }")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-D__GNUC_PYTHON__", f"-I{REF}/Include", f"-I{REF}/PrivateInclude", src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    return {k: int(v) for k, v in (line.split() for line in out.strip().splitlines())}


def static_assert_tu(layout):
    lines = ['#include <stddef.h>', '#include "arm_math.h"', '#include "arm_const_structs.h"']
    for key, v in sorted(layout.items()):
        parts = key.split("/")
        if parts[0] == "sizeof":
            lines.append(f'_Static_assert(sizeof({parts[1]}) == {v}, "sizeof({parts[1]}) differs from the reference");')
        elif parts[0] == "offsetof":
            lines.append(f'_Static_assert(offsetof({parts[1]}, {parts[2]}) == {v}, "offsetof({parts[1]}, {parts[2]}) differs from the reference");')
        else:
            lines.append(f'_Static_assert((int){parts[1]} == {v}, "{parts[1]} differs from the reference");')
    lines.append("int layout_checked;")
    return "\n".join(lines)


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not on this box (the committed layout file stands in)")
def test_committed_layout_file_is_what_the_reference_headers_say():
    assert probe_reference() == json.load(open(GOLDEN))


def test_repo_headers_static_assert_against_the_reference_layout():
    layout = json.load(open(GOLDEN))
    assert len(layout) == sum(1 + len(f) for f in STRUCTS.values()) + len(SCALARS) + len(STATUS)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "layout_tu.c")
        open(src, "w").write(static_assert_tu(layout))
        r = subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", f"-I{ROOT}/include", "-c", src, "-o", os.path.join(d, "layout_tu.o")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        # and the check has teeth: a wrong number does not compile
        bad = dict(layout)
        bad["sizeof/arm_rfft_fast_instance_f32"] += 8
        open(src, "w").write(static_assert_tu(bad))
        r = subprocess.run(["gcc", "-std=gnu11", f"-I{ROOT}/include", "-c", src, "-o", os.path.join(d, "bad.o")], capture_output=True, text=True)
        assert r.returncode != 0 and "differs from the reference" in r.stderr


if __name__ == "__main__" and "--write" in sys.argv:
    json.dump(probe_reference(), open(GOLDEN, "w"), indent=1, sort_keys=True)
    print("wrote", GOLDEN)
