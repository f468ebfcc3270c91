set -x
export CMSISDSP_CUDA_KERNEL=direct
CMD="python tools/sweep.py --mib 256 --reps 1 --warm 0 --ops rfft_fwd,rfft_inv,cfft_f32,cfft_q31,cfft_q15 --lens 1024,4096"
$CMD > gpurun_out/ncu1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -o gpurun_out/prof_direct $CMD > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
export CMSISDSP_CUDA_KERNEL=staged
CMD2="python tools/sweep.py --mib 256 --reps 1 --warm 0 --ops rfft_fwd,cfft_f32 --lens 1024,4096"
$CMD2 > gpurun_out/ncu2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_kernel -o gpurun_out/prof_staged $CMD2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
