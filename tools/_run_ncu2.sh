set -x
CMD="python tools/sweep.py --mib 256 --reps 1 --warm 0 --ops cfft_f32,rfft_fwd,rfft_inv --lens 512,1024,2048,4096"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -o gpurun_out/prof_v2 $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
