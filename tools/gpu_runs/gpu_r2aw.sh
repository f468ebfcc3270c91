#!/bin/bash
# round 2, call aw: L2 prefetch in the thread-per-frame kernel (N <= 64; real N <= 128)
{
for v in "" tpf1; do
  echo "== variant ${v:-default (no prefetch)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,cfft_mag,cfft_peak --lens 16,32,64 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfft_fwd,rfft_inv,rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 32,64,128 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2aw_tiny_prefetch.txt
