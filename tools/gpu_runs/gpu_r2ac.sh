#!/bin/bash
set -x
{
for rep in 1 2; do
for v in "" q10; do
  echo "== variant ${v:-default} (pass $rep)"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q15 --lens 128,512,1024,2048 2>&1 | grep "^cfft" | cut -c1-112
done
done
} | tee gpurun_out/r2ac_q15_minb10.txt
