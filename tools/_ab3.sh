CMSISDSP_CUDA_KERNEL=direct python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,rfft_fwd,rfft_inv > gpurun_out/ab_direct.txt 2>&1
CMSISDSP_CUDA_KERNEL=pipe python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f32,rfft_fwd,rfft_inv > gpurun_out/ab_pipe.txt 2>&1
paste -d"\n" gpurun_out/ab_direct.txt gpurun_out/ab_pipe.txt | cut -c1-125
