#!/bin/bash
set -x
{
for v in "" b3 b2; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_q15 --lens 4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2ad_4096_minb.txt
