#!/bin/bash
# A/B of library variants and env knobs (graph replay of hoic_step): tools/ab_r2b.sh  [obj] [n] [steps]
cd "$(dirname "$0")/.."
obj=${1:-box}; n=${2:-4096}; k=${3:-60}
run() { echo -n "$1 | "; shift; env "$@" timeout 200 python tools/gpu_time.py $obj $n $k graph 2>&1 | tail -1 | cut -c1-100; }
run base X=1
for v in variants/lib_*.so; do case $v in *stats*) continue;; esac; [ -f $v ] && run "$(basename $v)" HOIC_LIB=$PWD/$v; done
