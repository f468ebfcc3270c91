#!/bin/bash
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2g_pytest.txt
{
for v in "" m10 m12; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 256,512,1024,2048 2>&1 | grep "^cfft" | cut -c1-112
done
for v in "" m5 m6; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15,cfft_q31 --lens 4096 2>&1 | grep "^cfft" | cut -c1-112
done
} | tee gpurun_out/r2g_minb.txt
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112 | tee gpurun_out/r2g_rq31.txt
