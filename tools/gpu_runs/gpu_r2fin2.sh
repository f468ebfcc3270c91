#!/bin/bash
# multi-GPU: the C dispatcher over every visible device, the bench under torchrun
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_host_sanitizers.py -m gpu -x -q  2>&1 | tail -4 | tee gpurun_out/r2fin2_pytest.txt
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2fin2_bench_${N}gpu.json 2> gpurun_out/r2fin2_bench_${N}gpu.err
cut -c1-300 gpurun_out/r2fin2_bench_${N}gpu.json; tail -3 gpurun_out/r2fin2_bench_${N}gpu.err
