# f64 round: parity tests, smoke, bench, sweeps of the f64 entry points
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; cut -c1-330 gpurun_out/bench_last.json
