#!/bin/bash
# ncu full captures: two mid-circuit hea28 passes, and the persistent form on qft30
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 20 -c 2 -f -o gpurun_out/prof_tile_hea28 \
   python bench.py --workload hea28 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_hea28.log 2>&1
echo "ncu hea28 exit $?"
DVD_PERSIST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 3 -c 3 -f -o gpurun_out/prof_tile_qft30_persist \
   python bench.py --workload qft30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_qft30p.log 2>&1
echo "ncu qft30 persist exit $?"
ls -la gpurun_out/*.ncu-rep
