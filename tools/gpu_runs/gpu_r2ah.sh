#!/bin/bash
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "mfcc" 2>&1 | tail -2
{
for v in "" u2m6 u2m8 u4m4 u8m2; do
  echo "== mfcc variant ${v:-default (4 warps per CTA, 3 CTAs)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops mfcc --lens 256,512,1024 2>&1 | grep "^mfcc" | cut -c1-112
done
} | tee gpurun_out/r2ah_mfcc_geometry.txt
