#!/bin/bash
# round 2, call ap: MFCC stage ablation (timing only, results wrong): 1 = no mel loop, 6 = no partial sums / log / DCT, 7 = both
{
for v in "" abl1 abl6 abl7; do
  echo "== mfcc variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops mfcc --lens 256,512,1024 2>&1 | grep "^mfcc" | cut -c1-112
done
} | tee gpurun_out/r2ap_mfcc_ablation.txt
