#!/bin/bash
# round 2, call ar: L2 prefetch of the frame the slot's next CTA will load (direct kernels): distance 1x / 2x the resident CTAs
{
for v in "" pf1 pf2; do
  echo "== variant ${v:-default (no prefetch)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_q15,cfft_f64 --lens 256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv,rfftq15_inv --lens 512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2ar_prefetch.txt
