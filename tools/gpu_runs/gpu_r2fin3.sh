#!/bin/bash
# round 2, last call: the committed state once more -- GPU tests, smoke, both bench arms
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2fin3_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2fin3_smoke.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2fin3_bench_ref.json 2> gpurun_out/r2fin3_bench_ref.err; cut -c1-200 gpurun_out/r2fin3_bench_ref.json
python bench.py > gpurun_out/r2fin3_bench.json 2> gpurun_out/r2fin3_bench.err; cut -c1-300 gpurun_out/r2fin3_bench.json; tail -2 gpurun_out/r2fin3_bench.err
