#!/bin/bash
set -x
{
for v in "" occ4 occ2; do
  echo "== variant ${v:-default(occ3)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops rfft_fwd --lens 512,1024,2048 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2i_occ.txt
ncu --set full --clock-control none --import-source on -k regex:mfcc -s 3 -c 1 -o gpurun_out/r2i_prof_mfcc python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops mfcc --lens 1024 > gpurun_out/r2i_ncu_mfcc.log 2>&1; tail -1 gpurun_out/r2i_ncu_mfcc.log
