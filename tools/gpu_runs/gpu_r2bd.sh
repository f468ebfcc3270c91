#!/bin/bash
# round 2, call bd: fixed-point plan of N = 128 padded after every 8 elements (first-pass stores conflict-free): GPU tests, sweep of the units that use it
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2bd_pytest.txt
{
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q31,cfft_q15 --lens 128 2>&1 | grep "^cfft" | cut -c1-112
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 256 2>&1 | grep "^rfft" | cut -c1-112
} | tee gpurun_out/r2bd_fix128_pad.txt
