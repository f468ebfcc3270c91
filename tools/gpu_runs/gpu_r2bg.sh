#!/bin/bash
# round 2, call bg: BASELINE config 5 (arm_cfft_f32 length x batch sweep, CPU reference beside it) on one GPU with the end-of-round kernels; bench.py smoke of the edited MFCC row
set -x
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-sustained > gpurun_out/r2bg_bench_short.json 2> gpurun_out/r2bg_bench_short.err; tail -c 900 gpurun_out/r2bg_bench_short.json; tail -2 gpurun_out/r2bg_bench_short.err
( time python tools/sweep_config5.py --json gpurun_out/r2bg_config5_1gpu.json > gpurun_out/r2bg_config5_1gpu.txt 2>&1 ) 2>&1 | tail -3; tail -5 gpurun_out/r2bg_config5_1gpu.txt | cut -c1-200
