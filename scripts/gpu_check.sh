#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_qft30.log 2>&1; tail -2 gpurun_out/bench_qft30.log
timeout 300 python bench.py --workload hea28 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_hea28.log 2>&1; tail -1 gpurun_out/bench_hea28.log
timeout 300 python bench.py --workload qft30 --unfused --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_qft30_unfused.log 2>&1; tail -1 gpurun_out/bench_qft30_unfused.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_qft30.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
