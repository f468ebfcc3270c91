#!/bin/bash
set -x
{
for v in "" i4 i6 i7; do
  echo "== variant ${v:-default (KU_MINB 5)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv --lens 512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2aa_rifft_minb.txt
