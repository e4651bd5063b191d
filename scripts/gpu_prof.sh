#!/bin/bash
# microbench + remaining GPU tests + ncu full capture of the tile kernel on a 26-qubit QFT
mkdir -p gpurun_out
./scripts/microbench > gpurun_out/microbench.log 2>&1; cat gpurun_out/microbench.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 3 -c 3 -f -o gpurun_out/prof_tile_qft26 \
   python bench.py --workload qft26 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
