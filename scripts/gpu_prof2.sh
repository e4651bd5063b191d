#!/bin/bash
mkdir -p gpurun_out
W=${1:-qft30}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 3 -c 3 -f -o gpurun_out/prof_tile_${W}_v7 \
   python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_v7.log 2>&1
tail -2 gpurun_out/ncu_full_v7.log | cut -c1-200
