export CMSISDSP_CUDA_KERNEL=pipe
for len in 512 2048; do
CMD="python tools/sweep.py --mib 1024 --reps 1 --warm 1 --ops rfft_fwd --lens $len"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o gpurun_out/prof_v7_fwd$len $CMD >> gpurun_out/ncu.log 2>&1
done
tail -2 gpurun_out/ncu.log
