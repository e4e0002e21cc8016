#!/bin/bash
# compute-sanitizer over the final build of round 2 (run on the GPU box): default pipeline (memcheck, racecheck, synccheck), the
# sampler step's library kernels, and the opt-in schedules once under memcheck.  Output -> gpurun_out/r02b_sanitizer.txt
cd "$(dirname "$0")/.."
out=gpurun_out/r02b_sanitizer.txt
mkdir -p gpurun_out; : > $out
run() { echo "==== $*" >> $out; env "$@" 2>&1 | grep -v "^$" | tail -5 >> $out; }
for tool in memcheck racecheck synccheck; do
  run HOIC_SAN_RECORD=1 timeout 300 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py box 12 3
  run HOIC_SAN_RECORD=0 timeout 300 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py banana 12 1
done
run HOIC_SAN_RECORD=0 HOIC_STEP_PARTS=2 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py box 40 2
run HOIC_SAN_RECORD=1 HOIC_TICKS=96 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py box 12 1
run HOIC_SAN_RECORD=1 HOIC_SOLVE_BUDGET=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py box 12 1
run HOIC_SAN_RECORD=1 HOIC_FUSE12=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py box 12 1
run HOIC_SAN_SAMPLER=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py box 12 3
grep -E "^====|ERROR SUMMARY|RACECHECK SUMMARY|ok " $out
