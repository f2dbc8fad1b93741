#!/bin/sh
# Builds A/B variants of libeuler2d_b200.so into build/variants/<name>/ (development aid; `make` builds the product).
#   usage: tools/build_variants.sh name "-DFLAG=1 ..." [name "flags" ...]
# Run one with  E2D_LIB_PATH=build/variants/<name>/libeuler2d_b200.so python tools/quick_perf.py ...
set -e
cd "$(dirname "$0")/../euler2d_kokkos_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  out=../../build/variants/$name
  mkdir -p "$out"
  ( $NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v \
      $flags -shared -o "$out/libeuler2d_b200.so" e2d_kernels.cu e2d_slab.cu e2d_post.cu e2d_capi.cu e2d_config.cpp \
      2> "$out/ptxas.log" && echo "built $name" ) &
done
wait
