#!/bin/bash
# round 2, call aj: fixed-point N = 64 on four threads per frame (two passes, direct kernel) instead of one thread per frame (tiny kernel)
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "q15 or q31 or fixed or radix or boundary" 2>&1 | tail -2 | tee gpurun_out/r2aj_pytest.txt
{
for v in "" t1; do
  echo "== variant ${v:-four threads per frame}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_q15 --lens 64 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 128 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2aj_fix64.txt
