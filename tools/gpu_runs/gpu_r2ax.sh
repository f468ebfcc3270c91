#!/bin/bash
# round 2, call ax: tiny-kernel prefetch defaults applied (GPU tests + sweep of the short lengths); direct-kernel prefetch for the remaining f64 units
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2ax_pytest.txt
{
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,cfft_mag,cfft_peak --lens 16,32,64 2>&1 | grep "^cfft" | cut -c1-112
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq15_fwd,rfftq15_inv --lens 32,64,128 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2ax_tiny_prefetch_applied.txt
{
for v in "" pf1; do
  echo "== variant ${v:-default}"
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64 --lens 16,32,64,128,256 2>&1 | grep "^cfft" | cut -c1-112
  CMSISDSP_B200_LIBDIR=${v:+cmsis-dsp_b200/lib_$v} python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_inv,rfft64_fwd --lens 32,64,128,256,512,1024,2048,4096 2>&1 | grep "^rfft" | cut -c1-112
done
} | tee gpurun_out/r2ax_f64_prefetch.txt
