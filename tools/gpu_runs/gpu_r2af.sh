#!/bin/bash
set -x
{
for v in "" d5 d4; do
  echo "== variant ${v:-default}   (d5: q31 2048 MINB 5, rifft q31 NC 4096 MINB 3, f64 2048 MINB 6, rfft_q31 fwd NC 4096 MINB 2; d4: q31 2048 MINB 4, f64 2048 MINB 4, rfft_q31 fwd NC 4096 MINB 4)"
  L=${v:+cmsis-dsp_b200/lib_$v}
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops cfft_q31,cfft_f64 --lens 2048 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=$L python tools/sweep.py --mib 1024 --reps 30 --ops rfftq31_fwd,rfftq31_inv --lens 8192 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2af_minb.txt
