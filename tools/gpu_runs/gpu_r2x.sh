#!/bin/bash
# round 2, call x: units whose register count sits just above an occupancy step -- launch bounds one CTA tighter
set -x
{
for v in "" thr; do
  echo "== variant ${v:-default}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31 --lens 256,512,1024,2048 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 256 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_inv --lens 256,512 2>&1 | grep "^rfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd --lens 64,128 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2x_thresholds.txt
