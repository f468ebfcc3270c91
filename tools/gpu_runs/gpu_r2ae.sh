#!/bin/bash
set -x
{
for v in "" c8; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q15 --lens 128,512,2048 2>&1 | grep "^cfft" | cut -c1-112
done
for v in "" c3; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops rfftq15_fwd,rfftq15_inv --lens 8192 2>&1 | grep "^rfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops cfft_f64 --lens 4096 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_q15 --lens 4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2ae_minb.txt
