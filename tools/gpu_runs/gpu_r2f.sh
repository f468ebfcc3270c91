#!/bin/bash
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "rfft_fixed or rfft_fix" 2>&1 | tail -3
for v in "" rfnp rfm6 rfm5; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq15_fwd --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done | tee gpurun_out/r2f_rfix_prefetch.txt
