#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/opcost.py 28 > gpurun_out/opcost.log 2>&1; cat gpurun_out/opcost.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 3 -c 3 -f -o gpurun_out/prof_tile_qft26_v5 \
   python bench.py --workload qft26 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_v5.log 2>&1
tail -2 gpurun_out/ncu_full_v5.log | cut -c1-200
