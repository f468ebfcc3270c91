#!/bin/bash
set -x
{
for v in "" e32; do
  echo "== variant ${v:-default (16 points per thread)}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_q15 --lens 1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2am_fix_e32.txt
