for op in rfft_fwd rfft_inv cfft_f32; do
CMD="python tools/sweep.py --mib 1024 --reps 1 --warm 1 --ops $op --lens 4096"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v4_$op $CMD >> gpurun_out/ncu.log 2>&1
done
tail -3 gpurun_out/ncu.log
