#!/bin/bash
# round 2, call bb: q15 first-stage unpack shifts as IMAD.HI on the XU pipe (1: the y halves, 2: both halves)
{
for v in "" xu1 xu2; do
  echo "== variant ${v:-default (SHF on the ALU pipe)}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 128,256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2bb_q15_xu_unpack.txt
timeout 300 env CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_xu2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q15" 2>&1 | tail -3
