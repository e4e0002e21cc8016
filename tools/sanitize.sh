#!/bin/bash
# compute-sanitizer passes over the step pipeline (run on the GPU box): memcheck, racecheck, synccheck on a small batch in
# the grasp phase, box and banana, test-mode recording on and off.  Output -> gpurun_out/r02_sanitizer.txt
cd "$(dirname "$0")/.."
out=gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out; : > $out
for tool in memcheck racecheck synccheck; do
  for obj in box banana; do
    for rec in 1 0; do
      echo "==== compute-sanitizer --tool $tool   python tools/sanitize_small.py $obj 12 1 (record=$rec)" >> $out
      HOIC_SAN_RECORD=$rec timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py $obj 12 1 2>&1 | grep -v "^$" | tail -8 >> $out
    done
  done
done
grep -E "^====|ERROR SUMMARY|ok " $out
