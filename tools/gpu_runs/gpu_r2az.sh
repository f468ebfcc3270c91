#!/bin/bash
# round 2, call az: direct-kernel L2 prefetch for the remaining short fixed-point units (complex N = 64, 128; real N = 128, 256)
{
for v in "" pf1; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_q15 --lens 64,128 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 128,256 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2az_prefetch_short_fix.txt
