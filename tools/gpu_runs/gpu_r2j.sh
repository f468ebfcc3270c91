#!/bin/bash
set -x
{
for cap in 64 4 3 2; do
  echo "== CMSISDSP_CUDA_PIPE_MAXOCC=$cap"
  CMSISDSP_CUDA_PIPE_MAXOCC=$cap python tools/sweep.py --mib 1024 --reps 30 --ops cfft_f32,rfft_fwd,rfft_inv,cfft_mag,cfft_peak --lens 512,1024,2048,4096 2>&1 | grep -E "^(cfft|rfft)" | cut -c1-112
done
} | tee gpurun_out/r2j_pipe_occ.txt
