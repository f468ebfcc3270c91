#!/bin/bash
# One GPU round: parity tests, the bench (both arms), the ncu launch list and one full capture of the
# bench kernels, the throughput sweep of every entry point, ncu captures of the other kernel families.
# Run on a B200 box from the repo root:
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
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv,mfcc,cfft_mag,cfft_peak --json gpurun_out/sweep_g.json > gpurun_out/sweep_g.txt 2>&1; cat gpurun_out/sweep_g.txt | cut -c1-125
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 --json gpurun_out/sweep_rfix.json > gpurun_out/sweep_rfix.txt 2>&1; cat gpurun_out/sweep_rfix.txt | cut -c1-125
python tools/bench_config4.py --json gpurun_out/config4_1gpu.json > gpurun_out/config4_1gpu.txt 2>&1; tail -1 gpurun_out/config4_1gpu.txt | cut -c1-400
# one full capture per remaining kernel family (each after its command ran clean above)
for spec in "mfcc 1024 mfcc_kernel" "cfft_q31 1024 frame_kernel" "cfft_q15 1024 frame_kernel" "cfft_f32 1024 frame_kernel" "rfftq31_fwd 1024 frame_kernel" "cfft_peak 1024 frame_kernel"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/prof_fam_$1 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops $1 --lens $2 > gpurun_out/ncu_fam_$1.log 2>&1; tail -1 gpurun_out/ncu_fam_$1.log
done
