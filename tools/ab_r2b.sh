#!/bin/bash
cd "$(dirname "$0")/.."
obj=${1:-box}; n=${2:-4096}; k=${3:-60}
run() { echo -n "$1 | "; shift; env "$@" timeout 200 python tools/gpu_time.py $obj $n $k graph 2>&1 | tail -1 | cut -c1-230; }
run base X=1
run ticks64_b4 HOIC_TICKS=64 HOIC_TICK_BUDGET=4
run ticks64_b3 HOIC_TICKS=64 HOIC_TICK_BUDGET=3
run ticks56_b5 HOIC_TICKS=56 HOIC_TICK_BUDGET=5
run ticks80_b2 HOIC_TICKS=80 HOIC_TICK_BUDGET=2
