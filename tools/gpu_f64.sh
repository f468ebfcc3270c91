# f64 round: parity tests, smoke, bench, sweeps of the f64 entry points
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; cut -c1-330 gpurun_out/bench_f.json
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64,rfft64_fwd,rfft64_inv --json gpurun_out/sweep_f64_final.json > gpurun_out/sweep_f64_final.txt 2>&1; cut -c1-140 gpurun_out/sweep_f64_final.txt
