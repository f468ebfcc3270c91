#!/bin/bash
set -x
{
for v in "" q31m6; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31 --lens 512,1024,2048 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2w_q31_minb6.txt
