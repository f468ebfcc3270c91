#!/bin/bash
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "mfcc" 2>&1 | tail -3 | tee gpurun_out/r2s_pytest.txt
{
for v in "" mfccold; do
  echo "== mfcc variant ${v:-new}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops mfcc --lens 256,512,1024 2>&1 | grep "^mfcc" | cut -c1-112
done
} | tee gpurun_out/r2s_mfcc.txt
ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 2 -c 1 -o gpurun_out/r2s_prof_mfcc_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops mfcc --lens 1024 > gpurun_out/r2s_ncu_mfcc.log 2>&1; tail -1 gpurun_out/r2s_ncu_mfcc.log
