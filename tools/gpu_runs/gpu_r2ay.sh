#!/bin/bash
# round 2, call ay: pipelined f32 kernels with an L2 prefetch of the group after the one being copied in
{
for v in "" ppf; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_mag,cfft_peak --lens 512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfft_fwd,rfft_inv --lens 512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2ay_pipe_prefetch.txt
