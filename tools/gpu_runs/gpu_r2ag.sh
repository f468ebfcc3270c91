#!/bin/bash
set -x
{
for v in "" e3 e4 e5; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops rfft64_fwd,rfft64_inv --lens 512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
python tools/sweep.py --mib 1024 --reps 30 --ops cfft_f64 --lens 1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
python tools/sweep.py --mib 1024 --reps 30 --ops rfftq31_fwd,rfftq31_inv --lens 8192 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2ag_f64_minb.txt
