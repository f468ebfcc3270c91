CMD="python tools/sweep.py --mib 1024 --reps 1 --warm 1 --ops rfft_fwd --lens 4096"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v6_fwd $CMD >> gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
