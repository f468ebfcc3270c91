#!/bin/bash
# round 2, call a: parity after the addressing / q15 arithmetic rewrite, instruction probes, q15 variants, regression sweep
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
(cd tools/probes && ./probe_dpx) 2>&1 | tee gpurun_out/r2a_probe.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2a_pytest.txt
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 256,1024,4096 2>&1 | tee gpurun_out/r2a_q15_default.txt | cut -c1-130
for v in q15s0 q15s2 q15d0 q15s0d0; do
  CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_$v python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 --lens 256,1024,4096 2>&1 | tee gpurun_out/r2a_q15_$v.txt | cut -c1-130
done
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv,mfcc,cfft_mag,cfft_peak --json gpurun_out/r2a_sweep.json 2>&1 | tee gpurun_out/r2a_sweep.txt | cut -c1-130
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 2>&1 | tee gpurun_out/r2a_sweep_rfix.txt | cut -c1-130
python bench.py --steps 100 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-900 gpurun_out/r2a_bench.json
