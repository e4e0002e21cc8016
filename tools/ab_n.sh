#!/bin/bash
# A/B timing at a chosen size: tools/ab_n.sh <obj> <n_envs> name1 name2 ...
cd "$(dirname "$0")/.."
obj=$1; n=$2; shift; shift
for v in "$@"; do
  HOIC_LIB=$PWD/variants/lib_$v.so python tools/gpu_time.py $obj $n 10 2>&1 | grep "ms/step" | grep -v without | cut -c1-110
done
