#!/bin/bash
# round 2, final call: validation of the state to be judged -- GPU tests, smoke, both bench arms, ncu launch list and full captures of the bench command
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2fin_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2fin_smoke.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2fin_bench_ref.json 2> gpurun_out/r2fin_bench_ref.err; cut -c1-300 gpurun_out/r2fin_bench_ref.json
( time python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err ) 2> gpurun_out/r2fin_bench.time; cut -c1-700 gpurun_out/r2fin_bench.json; tail -3 gpurun_out/r2fin_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2fin_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --quick > gpurun_out/r2fin_ncu_launches.log 2>&1; tail -1 gpurun_out/r2fin_ncu_launches.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 6 -c 2 -o gpurun_out/r2fin_prof_bench python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --no-sustained > gpurun_out/r2fin_ncu_full.log 2>&1; tail -1 gpurun_out/r2fin_ncu_full.log | cut -c1-200
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv,mfcc,cfft_mag,cfft_peak --json gpurun_out/r2fin_sweep.json > gpurun_out/r2fin_sweep.txt 2>&1; cut -c1-125 gpurun_out/r2fin_sweep.txt | tail -80
python tools/sweep.py --mib 1024 --reps 20 --ops rfftq31_fwd,rfftq31_inv,rfftq15_fwd,rfftq15_inv --lens 32,64,128,256,512,1024,2048,4096,8192 > gpurun_out/r2fin_sweep_rfix.txt 2>&1; cut -c1-125 gpurun_out/r2fin_sweep_rfix.txt | tail -40
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64,rfft64_fwd,rfft64_inv --lens 16,32,64,128,256,512,1024,2048,4096 > gpurun_out/r2fin_sweep_f64.txt 2>&1; cut -c1-125 gpurun_out/r2fin_sweep_f64.txt | tail -30
