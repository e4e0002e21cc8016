#!/bin/bash
# A/B of the pipeline's occupancy knobs (extra dynamic shared memory per CTA lowers the resident CTAs per SM).
# usage: tools/ab_split.sh [obj] [n_envs]
obj=${1:-box}; n=${2:-4096}
for v in "" "HOIC_P3_PAD=4096" "HOIC_P3_PAD=8192" "HOIC_P3_PAD=12288" "HOIC_P3_PAD=20480" "HOIC_P1_PAD=4096" "HOIC_P1_PAD=12288" "HOIC_P2_PAD=4096" "HOIC_P2_PAD=12288" "HOIC_PT_PAD=8192"; do
  echo "== $v"
  env $v python tools/gpu_time.py $obj $n 20 2>&1 | head -1
done
