#!/bin/bash
# multi-GPU: the C dispatcher over every visible device, the bench under torchrun
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_boundary.py -m gpu -x -q -k "fan_out or several_host_threads" 2>&1 | tail -4 | tee gpurun_out/r2e_pytest.txt
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2e_bench_${N}gpu.json 2> gpurun_out/r2e_bench_${N}gpu.err
cut -c1-300 gpurun_out/r2e_bench_${N}gpu.json; tail -3 gpurun_out/r2e_bench_${N}gpu.err
python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2>&1; cut -c1-200 gpurun_out/r2e_bench_ref.json
