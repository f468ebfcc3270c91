#!/bin/bash
# One GPU round: parity tests, the bench (both arms), the ncu launch list and one full capture of the
# bench kernels, and the throughput sweep.  Run on a B200 box from the repo root:
#     gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
# then, back on the build box:  python tools/make_profiles.py <tag>   (writes profiles/<tag>_*)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 2600 gpurun_out/bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
CMD="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/bench_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 6 -c 2 -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv,mfcc --json gpurun_out/sweep_g.json > gpurun_out/sweep_g.txt 2>&1; cat gpurun_out/sweep_g.txt | cut -c1-125
