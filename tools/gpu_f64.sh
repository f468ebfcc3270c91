set -x
python -m pytest tests -m gpu -x -q -k "f64" 2>&1 | tail -5
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64 --json gpurun_out/sweep_f64.json > gpurun_out/sweep_f64.txt 2>&1; cat gpurun_out/sweep_f64.txt | cut -c1-140
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -12
