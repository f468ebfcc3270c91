#!/bin/bash
set -x
for spec in "rfftq31_inv 2048" "rfftq31_fwd 2048" "cfft_f64 4096"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2m_prof_$1_$2 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops $1 --lens $2 > gpurun_out/r2m_ncu_$1_$2.log 2>&1; tail -1 gpurun_out/r2m_ncu_$1_$2.log
done
