#!/bin/bash
set -x
{
for ex in 0 8192 16384 24576; do
  echo "== CMSISDSP_CUDA_DIRECT_EXTRA_SMEM=$ex"
  CMSISDSP_CUDA_DIRECT_EXTRA_SMEM=$ex python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_f64 --lens 512,1024,2048,4096 2>&1 | grep -E "^(cfft|rfft)" | cut -c1-112
  CMSISDSP_CUDA_DIRECT_EXTRA_SMEM=$ex python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv,rfft64_fwd,rfft64_inv --lens 1024,2048,4096 2>&1 | grep -E "^(cfft|rfft)" | cut -c1-112
done
} | tee gpurun_out/r2l_direct_occ.txt
