#!/bin/bash
# round 2, call bj: BASELINE config 4 (arm_mfcc_f32 front end, frames sharded over the GPUs of the box, CPU reference beside it), end-of-round kernels
N=$(nvidia-smi -L | wc -l)
if [ $N = 1 ]; then
  python tools/bench_config4.py --json gpurun_out/r2bj_config4_1gpu.json > gpurun_out/r2bj_config4_1gpu.txt 2>&1
  python tools/bench_config4.py --hop 160 --json gpurun_out/r2bj_config4_1gpu_hop160.json > gpurun_out/r2bj_config4_1gpu_hop160.txt 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/bench_config4.py --json gpurun_out/r2bj_config4_${N}gpu.json > gpurun_out/r2bj_config4_${N}gpu.txt 2>&1
fi
tail -n 2 gpurun_out/r2bj_config4_*gpu*.txt | cut -c1-500
