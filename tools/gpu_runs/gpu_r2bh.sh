#!/bin/bash
# round 2, call bh: BASELINE config 5 strong-scaled over all GPUs of the box, then the bench under torchrun (end-of-round kernels)
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/sweep_config5.py --json gpurun_out/r2bh_config5_${N}gpu.json > gpurun_out/r2bh_config5_${N}gpu.txt 2>&1; tail -4 gpurun_out/r2bh_config5_${N}gpu.txt | cut -c1-160
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2bh_bench_${N}gpu.json 2> gpurun_out/r2bh_bench_${N}gpu.err; tail -n 1 gpurun_out/r2bh_bench_${N}gpu.json | cut -c1-260
