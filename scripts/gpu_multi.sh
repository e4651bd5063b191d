#!/bin/bash
# multi-GPU run (gpurun --gpus N): parity vs oracle on N ranks, then the strong-scaling bench (cfg 4) and,
# at N=8, the 34-qubit config (cfg 5).  Outputs under gpurun_out/.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_N=${DIST_CHECK_N:-16,24} timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
grep -E "^n=|DIST_CHECK" gpurun_out/dist_check_${N}gpu.log | tail -12
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_random32_${N}gpu.log 2>&1
tail -1 gpurun_out/bench_random32_${N}gpu.log | cut -c1-600
if [ "$N" = "8" ]; then
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload hea34 --steps 2 --warmup 1 > gpurun_out/bench_hea34_${N}gpu.log 2>&1
tail -1 gpurun_out/bench_hea34_${N}gpu.log | cut -c1-600
fi
