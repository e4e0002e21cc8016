#!/bin/bash
cd "$(dirname "$0")/.."
obj=${1:-box}; n=${2:-4096}; k=${3:-60}
run() { echo -n "$1 | "; shift; env "$@" timeout 200 python tools/gpu_time.py $obj $n $k graph 2>&1 | tail -1 | cut -c1-100; }
run base X=1
for bgt in 3 4 6; do run budget$bgt HOIC_SOLVE_BUDGET=$bgt; done
run budget4_prio0 HOIC_SOLVE_BUDGET=4 HOIC_PRIO_CARRY=0
run budget4_auxhi HOIC_SOLVE_BUDGET=4 HOIC_PRIO_AUX=-100
run budget4_1part HOIC_SOLVE_BUDGET=4 HOIC_STEP_PARTS=1
run base_1part HOIC_STEP_PARTS=1
