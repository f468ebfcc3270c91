set -x
python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64 --lens 1024,2048,4096 > gpurun_out/sweep_f64_b.txt 2>&1; cut -c1-140 gpurun_out/sweep_f64_b.txt
CMSISDSP_B200_LIBDIR=$PWD/cmsis-dsp_b200/lib_a python tools/sweep.py --mib 1024 --reps 20 --ops cfft_f64 --lens 1024,2048,4096 > gpurun_out/sweep_f64_a.txt 2>&1; cut -c1-140 gpurun_out/sweep_f64_a.txt
CMSISDSP_B200_LIBDIR=$PWD/cmsis-dsp_b200/lib_a python -m pytest tests -m gpu -x -q -k "f64" 2>&1 | tail -2
