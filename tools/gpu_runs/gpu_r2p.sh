#!/bin/bash
# e2e staging under contention: all GPUs of the box streaming at once
set -x
N=$(nvidia-smi -L | wc -l)
for cfg in "32 3" "32 6" "64 6" "16 6" "128 4" "8 8"; do
  set -- $cfg
  CMSISDSP_CUDA_CHUNK_MIB=$1 CMSISDSP_CUDA_NSTREAMS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-secondary --no-cpu --no-sustained > gpurun_out/r2p_e2e_$1_$2.json 2> gpurun_out/r2p_e2e_$1_$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_e2e_$1_$2.json").read().strip().splitlines()[-1])
e=d["e2e"]; o=e.get("one_process_all_devices",{})
print("chunk $1 MiB streams $2: e2e", round(e["value"]), "Msamples/s,", round(e["pcie_gbs_per_direction"],2), "GB/s per direction per rank, copy peak both", round(e["pinned_copy_peak"]["both_gbs_per_direction"],2), "frac", round(e["frac_of_pinned_copy_peak"],3), "| one process:", round(o.get("pcie_gbs_per_direction_total",0),1), "GB/s total")
PY
done 2>&1 | grep -v "^+" | tee gpurun_out/r2p_e2e_staging_8gpu.txt
