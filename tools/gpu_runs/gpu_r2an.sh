#!/bin/bash
# round 2, call an: chunk-size ramp of the host staging pipeline (first / last chunks small): boundary tests, sanitizer run, sweep
set -x
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_host_sanitizers.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2an_pytest.txt
python tools/sweep_staging.py --chunks 16,32,64,128 --streams 3 --ramps 0,1,2,4,8 --json gpurun_out/r2an_staging.json 2>&1 | tee gpurun_out/r2an_staging.txt
python tools/sweep_staging.py --chunks 32,64 --streams 2,4 --ramps 0,4 2>&1 | tee gpurun_out/r2an_staging_b.txt
