#!/bin/bash
# round 2, call q: rounding multiply-accumulates as IMAD.HI (XU) + IADD3 -- parity, A/B against the 64-bit form and launch bounds; MFCC ncu capture
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2q_pytest.txt
{
for v in "" wide m0 m5; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv,rfftq31_fwd --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done
for v in "" wide; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31 --lens 32,128,512,2048 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2q_rmac.txt
ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 2 -c 1 -o gpurun_out/r2q_prof_mfcc_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops mfcc --lens 1024 > gpurun_out/r2q_ncu_mfcc.log 2>&1; tail -1 gpurun_out/r2q_ncu_mfcc.log
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2q_prof_rfftq31_inv_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops rfftq31_inv --lens 1024 > gpurun_out/r2q_ncu_rifft.log 2>&1; tail -1 gpurun_out/r2q_ncu_rifft.log
