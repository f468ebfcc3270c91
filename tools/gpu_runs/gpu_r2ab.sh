#!/bin/bash
# round 2, call ab: launch-bound sweep of the fixed-point real FFT units (and a check of the rfft_q31 inverse after its change)
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "rfft_fixed or rfft_q or rfix" 2>&1 | tail -2 | tee gpurun_out/r2ab_pytest.txt
{
echo "== default"
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_inv,rfftq31_fwd,rfftq15_fwd,rfftq15_inv --lens 256,512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
for v in v6 v7 v8; do
  echo "== variant $v (rfft_q31 forward)"
  CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_$v python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd --lens 256,512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
for v in v7 v8; do
  echo "== variant $v (rfft_q15 forward)"
  CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_$v python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd --lens 256,512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
echo "== variant v10 (rfft_q15 inverse)"
CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_v10 python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_inv --lens 256,512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2ab_rfix_minb.txt
