#!/bin/bash
set -x
{
echo "== default caps"
python tools/sweep.py --mib 1024 --reps 30 --ops cfft_f32,rfft_fwd,rfft_inv,cfft_mag,cfft_peak --lens 512,1024,2048,4096 2>&1 | grep -E "^(cfft|rfft)" | cut -c1-112
echo "== CMSISDSP_CUDA_PIPE_MAXOCC=1"
CMSISDSP_CUDA_PIPE_MAXOCC=1 python tools/sweep.py --mib 1024 --reps 30 --ops cfft_f32,rfft_fwd,rfft_inv --lens 512,1024,2048,4096 2>&1 | grep -E "^(cfft|rfft)" | cut -c1-112
} | tee gpurun_out/r2k_caps.txt
{
for v in "" mq2 mq1; do
  echo "== mfcc variant ${v:-default(quad 4)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops mfcc --lens 256,512,1024 2>&1 | grep -E "^mfcc" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mfcc 2>&1 | tail -1
done
} | tee gpurun_out/r2k_mfcc.txt
