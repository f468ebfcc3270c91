# f64 round: parity tests, smoke, f64 sweeps
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -4
python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_fwd,rfft64_inv --lens 32,256,1024,4096 > gpurun_out/sweep_rfft64.txt 2>&1; cut -c1-140 gpurun_out/sweep_rfft64.txt
