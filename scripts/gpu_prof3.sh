#!/bin/bash
# v10 profiles: launch list (time + DRAM bytes) of the default bench, full capture of the three dense passes at 30 qubits
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_qft30.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-point > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 6 -c 3 -f -o gpurun_out/prof_tile_qft30_v10 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-point --no-single-gate > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep | tail -2
