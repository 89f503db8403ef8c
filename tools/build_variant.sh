#!/bin/bash
# Builds a variant of the library with extra nvcc flags: tools/build_variant.sh NAME "-DFR_PF=2 -DFR_POLL_NS=300"
# -> framefusion_b200/variants/libff_NAME.so (git-ignored; select it with FF_LIB_PATH).  Measurement aid.
set -e
cd "$(dirname "$0")/.."
mkdir -p framefusion_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $2 -I include \
  -o framefusion_b200/variants/libff_$1.so framefusion_b200/csrc/ff_api.cu
echo framefusion_b200/variants/libff_$1.so
