#!/bin/bash
# round 2, call bc: why is the fixed-point forward real FFT slow at real N = 256 (complex 128, 8 threads per frame)?  ncu of q31 and q15
for k in q31 q15; do
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/r2bc_prof_rfft${k}_fwd_256 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops rfft${k}_fwd --lens 256 > gpurun_out/r2bc_ncu_$k.log 2>&1; tail -1 gpurun_out/r2bc_ncu_$k.log
done
