# f64 round: parity tests, bench (regression check of the bench kernels after the header changes), f64 sweep,
# ncu launch list + full captures of the f64 kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 2300 gpurun_out/bench_f.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_f_ref.json 2> gpurun_out/bench_f_ref.err; cut -c1-300 gpurun_out/bench_f_ref.json
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64 --json gpurun_out/sweep_f64.json > gpurun_out/sweep_f64.txt 2>&1; cut -c1-140 gpurun_out/sweep_f64.txt
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv --lens 1024,4096 > gpurun_out/sweep_regress.txt 2>&1; cut -c1-140 gpurun_out/sweep_regress.txt
for n in 1024 4096; do
  ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/prof_f64_$n python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops cfft_f64 --lens $n > gpurun_out/ncu_f64_$n.log 2>&1; tail -1 gpurun_out/ncu_f64_$n.log
done
