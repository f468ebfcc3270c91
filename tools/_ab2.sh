OPS="${OPS:-cfft_f32,rfft_fwd,rfft_inv}"; LENS="${LENS:-2048,4096}"
for fl in 0 1 2 3; do
echo "== flags $fl"
CMSISDSP_CUDA_PIPE_FLAGS=$fl CMSISDSP_CUDA_KERNEL=pipe python tools/sweep.py --mib 1024 --reps 20 --ops $OPS --lens $LENS 2>&1 | cut -c1-120
done
