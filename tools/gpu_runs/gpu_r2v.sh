#!/bin/bash
# round 2, call v: q31 adds pinned to the ALU / FMA pipes -- parity, A/B against ptxas' own placement
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "q31 or fixed or radix or boundary" 2>&1 | tail -3 | tee gpurun_out/r2v_pytest.txt
{
for v in "" noplace; do
  echo "== variant ${v:-placed}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2v_q31_place.txt
