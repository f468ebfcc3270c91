#!/bin/bash
set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "rfft or fixed or boundary" 2>&1 | tail -2 | tee gpurun_out/r2ai_pytest.txt
python tools/sweep.py --mib 1024 --reps 30 --ops rfftq15_inv,rfftq31_inv --lens 64,128,256,512,1024,2048,4096,8192 2>&1 | grep "^rfft" | cut -c1-112 | tee gpurun_out/r2ai_rifft_q15_pack.txt
