#!/bin/bash
# Round-2 profile captures (run on the GPU box); outputs under gpurun_out/, summarised into profiles/ by tools/summarise_profiles.py
#  1. launch list of the bench command (ncu --metrics gpu__time_duration.sum, graph kernel nodes profiled individually)
#  2. DRAM traffic + durations of every launch of a few hoic_step calls (box / 4096 envs, default 2 parts)
#  3. ncu --set full of one launch of each pipeline kernel at box / 4096 envs (HOIC_STEP_PARTS=1: whole-batch kernels)
cd "$(dirname "$0")/.."
TAG=${TAG:-r02_final}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --preroll 2 --no-extras --no-cpu > gpurun_out/${TAG}_bench_under_ncu.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 2600 --launch-count 520 --csv \
  --log-file gpurun_out/${TAG}_step_traffic.csv python tools/gpu_time.py box 4096 8 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --launch-skip 2600 --launch-count 520 --csv \
  --log-file gpurun_out/${TAG}_step_traffic_warm.csv python tools/gpu_time.py box 4096 8 > /dev/null 2>&1
for kn in k_p1 k_p2 k_p3s k_tail; do
  HOIC_STEP_PARTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$kn\$ --launch-skip 200 --launch-count 1 \
    -o gpurun_out/${TAG}_$kn -f python tools/gpu_time.py box 4096 3 > /dev/null 2>&1
done
HOIC_STEP_PARTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_p2\$ --launch-skip 200 --launch-count 1 \
  -o gpurun_out/${TAG}_banana_k_p2 -f python tools/gpu_time.py banana 4096 3 > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
