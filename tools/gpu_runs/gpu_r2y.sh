#!/bin/bash
# round 2, call y: q15 raw exchange (upper halves loaded with LDS.U16 / S16 instead of shifting in the producer) -- parity, A/B
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "q15 or fixed or radix or boundary" 2>&1 | tail -3 | tee gpurun_out/r2y_pytest.txt
{
for v in "" noraw; do
  echo "== variant ${v:-raw exchange}"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 128,256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd,rfftq15_inv --lens 512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2y_q15_raw.txt
