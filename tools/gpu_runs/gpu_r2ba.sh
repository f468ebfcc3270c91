#!/bin/bash
# round 2, call ba: direct kernel, a CTA runs 2 / 4 frame groups in sequence (set-up and CTA launch amortised), no loads in flight across groups
{
for v in "" loop2 loop4; do
  echo "== variant ${v:-default (one group per CTA)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_q15 --lens 256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2ba_loop.txt
timeout 300 env CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_loop2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfft_all_modes or config3" 2>&1 | tail -3
