set -x
CMSISDSP_B200_LIBDIR=$PWD/cmsis-dsp_b200/lib_b python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_fwd,rfft64_inv --lens 32,64,128,256,512,1024,2048,4096 > gpurun_out/sweep_rfft64_minb1.txt 2>&1; cut -c1-140 gpurun_out/sweep_rfft64_minb1.txt
