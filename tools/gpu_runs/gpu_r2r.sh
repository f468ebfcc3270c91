#!/bin/bash
# round 2, call r: MFCC mel stage rewritten (row partial sums, no flush) -- parity, A/B against the old library, ncu; rfft_q31 inverse with 5 CTAs
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "mfcc or rfft or boundary" 2>&1 | tail -5 | tee gpurun_out/r2r_pytest.txt
{
for v in "" mfccold; do
  echo "== mfcc variant ${v:-new}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops mfcc --lens 256,512,1024 2>&1 | grep "^mfcc" | cut -c1-112
done
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2r_mfcc.txt
ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 2 -c 1 -o gpurun_out/r2r_prof_mfcc_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops mfcc --lens 1024 > gpurun_out/r2r_ncu_mfcc.log 2>&1; tail -1 gpurun_out/r2r_ncu_mfcc.log
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2r_prof_rfftq31_inv_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops rfftq31_inv --lens 1024 > gpurun_out/r2r_ncu_rifft.log 2>&1; tail -1 gpurun_out/r2r_ncu_rifft.log
