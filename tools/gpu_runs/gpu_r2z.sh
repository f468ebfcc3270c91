#!/bin/bash
# round 2, call z (final state): validation of the state to be judged -- GPU tests, smoke, both bench arms, ncu launch list and full captures of the bench command
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2zz_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2zz_smoke.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2zz_bench_ref.json 2> gpurun_out/r2zz_bench_ref.err; cut -c1-300 gpurun_out/r2zz_bench_ref.json
python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; cut -c1-700 gpurun_out/r2zz_bench.json; tail -3 gpurun_out/r2zz_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --quick > gpurun_out/r2zz_ncu_launches.log 2>&1; tail -1 gpurun_out/r2zz_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 6 -c 2 -o gpurun_out/r2zz_prof_bench python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --no-sustained > gpurun_out/r2zz_ncu_full.log 2>&1; tail -1 gpurun_out/r2zz_ncu_full.log | cut -c1-200
