#!/bin/bash
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "q31 or q15 or fixed" 2>&1 | tail -2
{
for v in "" f1; do
  echo "== variant ${v:-default (two frames per CTA at N = 1024)}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_q15 --lens 64,1024 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 2048 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2ak_fix1024_f1.txt
