#!/bin/bash
# round 2, call bf: ncu of the units that sit below their neighbours in the sweep
for spec in "rfft_fwd 2048" "cfft_q31 64" "rfftq31_fwd 128" "rfftq15_fwd 256" "rfftq15_inv 1024" "rfftq31_inv 1024" "cfft_q15 16"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2bf_prof_$1_$2 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops $1 --lens $2 > gpurun_out/r2bf_ncu_$1_$2.log 2>&1; tail -1 gpurun_out/r2bf_ncu_$1_$2.log
done
