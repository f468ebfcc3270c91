#!/bin/bash
# round 2, call aq: ncu of the MFCC front half alone (variant abl7: no mel loop / log / DCT) and of cfft_peak N = 512 for comparison
set -x
CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_abl7 ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 2 -c 1 -o gpurun_out/r2aq_prof_mfcc_front python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops mfcc --lens 1024 > gpurun_out/r2aq_ncu_mfcc.log 2>&1; tail -1 gpurun_out/r2aq_ncu_mfcc.log
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/r2aq_prof_peak_512 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops cfft_peak --lens 512 > gpurun_out/r2aq_ncu_peak.log 2>&1; tail -1 gpurun_out/r2aq_ncu_peak.log
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/r2aq_prof_rfft_1024 python tools/sweep.py --mib 256 --reps 2 --warm 2 --ops rfft_fwd --lens 1024 > gpurun_out/r2aq_ncu_rfft.log 2>&1; tail -1 gpurun_out/r2aq_ncu_rfft.log
