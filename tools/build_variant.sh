#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...]: builds build_variants/libsdrgpu_NAME.so (kernel-variant experiments;
# select it at run time with SDRGPU_LIB=build_variants/libsdrgpu_NAME.so).  build_variants/ is git-ignored.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o build_variants/libsdrgpu_$name.so sdrainer_b200/csrc/engine.cu sdrainer_b200/csrc/goertzel.cu
echo built build_variants/libsdrgpu_$name.so
