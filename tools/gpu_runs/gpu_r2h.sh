#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2h_pytest.txt
{
for fl in direct pipe; do
  echo "== CMSISDSP_CUDA_KERNEL=$fl"
  CMSISDSP_CUDA_KERNEL=$fl python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15,cfft_q31 --lens 128,256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2h_loop_vs_direct.txt
