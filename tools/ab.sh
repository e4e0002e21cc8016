#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh obj n_envs name1 name2 ...   (variants/lib_<name>.so)
cd "$(dirname "$0")/.."
obj=$1; n=$2; shift; shift
for v in "$@"; do
  echo -n "$v: "; HOIC_LIB=$PWD/variants/lib_$v.so python tools/gpu_time.py $obj $n 20 2>&1 | tail -1
done
