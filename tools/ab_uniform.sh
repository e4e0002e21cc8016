#!/bin/bash
# uniform-env A/B: tools/ab_uniform.sh "<starts>" name1 name2 ...  ("default" = the in-tree library)
cd "$(dirname "$0")/.."
starts=$1; shift
for v in "$@"; do
  for st in $starts; do
    if [ "$v" = default ]; then python tools/gpu_time_uniform.py box 4096 10 $st 2>&1 | tail -1
    else HOIC_LIB=$PWD/variants/lib_$v.so python tools/gpu_time_uniform.py box 4096 10 $st 2>&1 | tail -1; fi
  done
done
