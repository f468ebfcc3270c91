# f64 round: parity tests, smoke, rfft f64 chunk-size sweep
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
for c in 0 4 8 16 32 64; do
  echo "## CMSISDSP_CUDA_RFFT64_CHUNK_MIB=$c" >> gpurun_out/sweep_rfft64_chunks.txt
  CMSISDSP_CUDA_RFFT64_CHUNK_MIB=$c python tools/sweep.py --mib 1024 --reps 20 --ops rfft64_fwd,rfft64_inv --lens 32,256,1024,4096 >> gpurun_out/sweep_rfft64_chunks.txt 2>&1
done
cut -c1-110 gpurun_out/sweep_rfft64_chunks.txt
