#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh name1 name2 ...   (variants/lib_<name>.so)
cd "$(dirname "$0")/.."
for v in "$@"; do
  HOIC_LIB=$PWD/variants/lib_$v.so python tools/gpu_time.py box 4096 20 2>&1 | tail -2
done
