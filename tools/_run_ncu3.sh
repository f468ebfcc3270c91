CMD="python tools/sweep.py --mib 512 --reps 1 --warm 1 --ops cfft_f32,rfft_fwd,rfft_inv --lens 4096"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v3a $CMD > gpurun_out/ncu.log 2>&1
CMD="python tools/sweep.py --mib 512 --reps 1 --warm 1 --ops rfft_fwd --lens 4096"
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v3b $CMD >> gpurun_out/ncu.log 2>&1
CMD="python tools/sweep.py --mib 512 --reps 1 --warm 1 --ops rfft_inv --lens 4096"
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v3c $CMD >> gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
