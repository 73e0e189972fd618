#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...] -> build/variants/libakari_b200_NAME.so (kernel-tuning experiments)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -std=c++17 -O3 -prec-div=false -prec-sqrt=false -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o build/variants/libakari_b200_$name.so akari_render_b200/csrc/akari_b200.cu akari_render_b200/csrc/host/scene_build.cpp
