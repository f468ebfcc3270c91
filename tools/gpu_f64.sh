# f64 round: parity tests, smoke, sweeps of the f64 entry points, one ncu capture of the fused rfft f64 kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_fwd,rfft64_inv --lens 32,64,128,256,512,1024,2048,4096 > gpurun_out/sweep_rfft64_fused.txt 2>&1; cut -c1-140 gpurun_out/sweep_rfft64_fused.txt
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/prof_rfft64_fwd_4096 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops rfft64_fwd --lens 4096 > gpurun_out/ncu_rfft64.log 2>&1; tail -1 gpurun_out/ncu_rfft64.log
