#!/bin/bash
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "q15 or fixed" 2>&1 | tail -2 | tee gpurun_out/r2al_pytest.txt
{
for v in "" g18 g20; do
  echo "== variant ${v:-default}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q15 --lens 1024 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops rfftq15_fwd --lens 2048 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2al_q15_f1_minb.txt
