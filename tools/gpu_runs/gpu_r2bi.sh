#!/bin/bash
# round 2, call bi: BASELINE config 5 strong-scaled over the GPUs of the box (2 / 4)
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/sweep_config5.py --json gpurun_out/r2bi_config5_${N}gpu.json > gpurun_out/r2bi_config5_${N}gpu.txt 2>&1; tail -3 gpurun_out/r2bi_config5_${N}gpu.txt | cut -c1-160
