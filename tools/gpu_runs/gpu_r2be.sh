#!/bin/bash
# round 2, call be: arm_rfft_fast_f32 inverse N = 256 with the exchange padded after every 8 elements; GPU tests
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2be_pytest.txt
python tools/sweep.py --mib 1024 --reps 20 --ops rfft_fwd,rfft_inv --lens 128,256,512 2>&1 | grep "^rfft" | cut -c1-112 | tee gpurun_out/r2be_rfft_inv_256.txt
