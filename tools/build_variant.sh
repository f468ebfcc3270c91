#!/bin/bash
# A/B builds of compile-time choices: rebuild the kernel units of some ops with extra flags into
# cmsis-dsp_b200/lib_<name>/ (select it with CMSISDSP_B200_LIBDIR=cmsis-dsp_b200/lib_<name>).
#   tools/build_variant.sh <name> "<extra nvcc flags>" "<ops>" ["<lengths>"]
# e.g. tools/build_variant.sh q15alu "-DFFT_Q15_SHIFT_MODE=0" "2 7 8"
set -e
name=$1; flags=$2; ops=$3; lens=${4:-"16 32 64 128 256 512 1024 2048 4096"}
cd "$(dirname "$0")/../cmsis-dsp_b200/csrc"
B=../build/var_$name; L=../lib_$name
mkdir -p $B $L
cp -u ../build/*.o $B/ 2>/dev/null || true
for op in $ops; do
  if [ "$op" = mfcc ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -I../../include -Icuda $flags -c cuda/mfcc_unit.cu -o $B/mfcc.o &
    continue
  fi
  for n in $lens; do
  [ -f ../build/ku_${op}_${n}.o ] || continue
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -Xptxas -v -I../../include -Icuda \
      -DKU_OP=$op -DKU_N=$n $flags -c cuda/kernel_unit.cu -o $B/ku_${op}_${n}.o 2> $B/ku_${op}_${n}.ptxas || { tail -5 $B/ku_${op}_${n}.ptxas; exit 1; } ) &
  while [ $(jobs -r | wc -l) -ge 8 ]; do wait -n; done
done; done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a $B/shim.o $B/mfcc.o $B/radix2_fix.o $B/ku_*.o -o $L/libcmsisdsp_cuda.so
cp ../lib/libcmsisdsp_b200.so $L/      # rpath $ORIGIN: picks the variant's shim next to it
echo "built $L"
