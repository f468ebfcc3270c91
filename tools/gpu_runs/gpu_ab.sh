set -x
python -m pytest tests -m gpu -x -q -k "f64" 2>&1 | tail -2
python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_fwd --lens 32,64,128,256,512,1024,2048,4096 > gpurun_out/sweep_rfft64_fwd_prefetch.txt 2>&1; cut -c1-140 gpurun_out/sweep_rfft64_fwd_prefetch.txt
