#!/bin/bash
# v11 (run-time specialised kernels): launch list of the default bench and a full capture of the three dense passes
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_qft30_v11.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-point --no-single-gate > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dvd_pass_static -s 9 -c 3 -f -o gpurun_out/prof_jit_qft30_v11 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-point --no-single-gate > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/prof_jit_qft30_v11.ncu-rep
