#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2d_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2d_smoke.txt
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 2>&1 | tee gpurun_out/r2d_sweep_rfix.txt | cut -c1-130
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64,rfft64_fwd,rfft64_inv 2>&1 | tee gpurun_out/r2d_sweep_f64.txt | cut -c1-130
