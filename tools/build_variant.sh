#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]  -> variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --fmad=true \
  -prec-div=false -prec-sqrt=false -ftz=true "$@" -o variants/lib_$name.so hoic_b200/csrc/hoic_capi.cu
echo variants/lib_$name.so
