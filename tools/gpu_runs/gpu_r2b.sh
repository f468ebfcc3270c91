#!/bin/bash
# round 2, call b: parity incl. the new boundary tests, q15 after the biased-saturation rewrite (+ the FMA-adds variant)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.txt
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 2>&1 | tee gpurun_out/r2b_q15_default.txt | cut -c1-130
CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_q15fa python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15 2>&1 | tee gpurun_out/r2b_q15_fa.txt | cut -c1-130
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 2>&1 | tee gpurun_out/r2b_rq15_default.txt | cut -c1-130
CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_q15fa python tools/sweep.py --mib 1024 --reps 20 --ops rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 2>&1 | tee gpurun_out/r2b_rq15_fa.txt | cut -c1-130
python bench.py --steps 20 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cut -c1-400 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
