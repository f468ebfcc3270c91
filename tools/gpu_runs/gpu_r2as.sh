#!/bin/bash
# round 2, call as: L2 prefetch (distance = resident CTAs) for the other direct kernels: fixed-point forward real FFT, f64 real FFT, short f32 lengths
{
for v in "" pf1; do
  echo "== variant ${v:-default (no prefetch)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq15_fwd,rfft64_fwd,rfft64_inv --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_mag,cfft_peak,rfft_fwd,rfft_inv --lens 128,256 2>&1 | grep "^cfft\|^rfft" | cut -c1-112
done
} | tee gpurun_out/r2as_prefetch_b.txt
