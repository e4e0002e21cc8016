#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]  -> variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
HOIC_BUILD_OUT=$PWD/variants/lib_$name.so HOIC_NVCC_FLAGS="$*" python hoic_b200/build.py --force
rm -rf variants/lib_$name.so.obj
