python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMSISDSP_CUDA_KERNEL=direct python -m pytest tests -m gpu -x -q -k "rfft or cfft_all" 2>&1 | tail -3
OPS="${OPS:-cfft_f32,rfft_fwd,rfft_inv}"; LENS="${LENS:-256,512,1024,2048,4096}"
CMSISDSP_CUDA_KERNEL=direct python tools/sweep.py --mib 1024 --reps 20 --ops $OPS --lens $LENS > gpurun_out/ab_direct.txt 2>&1
CMSISDSP_CUDA_KERNEL=pipe python tools/sweep.py --mib 1024 --reps 20 --ops $OPS --lens $LENS > gpurun_out/ab_pipe.txt 2>&1
paste -d"\n" gpurun_out/ab_direct.txt gpurun_out/ab_pipe.txt
