#!/bin/bash
# round 2, call c: parity, the new bench line (both arms), q15 sweep, ncu of the q15 and bench kernels, staging sweep
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.txt
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-600 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; cut -c1-300 gpurun_out/r2c_bench_ref.json
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_q15,cfft_q31 2>&1 | tee gpurun_out/r2c_sweep_fix.txt | cut -c1-130
python tools/sweep_staging.py --json gpurun_out/r2c_staging.json 2>&1 | tee gpurun_out/r2c_staging.txt
for spec in "cfft_q15 1024" "cfft_q15 4096" "cfft_q31 4096" "rfft_fwd 4096" "rfft_inv 4096"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2c_prof_$1_$2 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops $1 --lens $2 > gpurun_out/r2c_ncu_$1_$2.log 2>&1; tail -1 gpurun_out/r2c_ncu_$1_$2.log
done
