#!/bin/bash
# Build A/B variants of libgh_cuda.so (same sources, different -D flags) into crime_b200/csrc/variants/;
# tools/ab_stage.py picks one with GH_CUDA_LIB=<path>.   usage: tools/build_variants.sh name "-DFLAG=.. -DFLAG2=.."  [...]
set -e
cd "$(dirname "$0")/../crime_b200/csrc"
mkdir -p variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVF="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=variants/$name; mkdir -p $d
  for f in gh_api gh_fft gh_kgen gh_fields gh_joint gh_psources; do nvcc $NVF $flags -c $f.cu -o $d/$f.o & done
  nvcc $NVF -fmad=false $flags -c gh_pixelize.cu -o $d/gh_pixelize.o &
  wait
  nvcc $ARCH -shared -o variants/libgh_cuda_$name.so $d/*.o -lnccl
  rm -rf $d
  echo built variants/libgh_cuda_$name.so
done
