#!/bin/bash
# round 2, call at: per-unit L2 prefetch defaults applied -- full GPU test suite and the sweep of the units that changed
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2at_pytest.txt
{
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_f64 --lens 256,512,1024,2048,4096 2>&1 | grep "^cfft" | cut -c1-112
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfft64_fwd --lens 256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2at_prefetch_applied.txt
