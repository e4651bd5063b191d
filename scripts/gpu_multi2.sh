#!/bin/bash
# N-GPU validation of the peer-memory swap path: parity vs oracle, then cfg 4 with the direct NVLink kernel
# and with the staged NCCL send/recv path (DVD_SWAP=nccl) for comparison.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_N=${DIST_CHECK_N:-16,22} timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
grep -E "^n=|DIST_CHECK|Error|error" gpurun_out/dist_check_${N}gpu.log | tail -12
for mode in peer nccl; do
  DVD_SWAP=$mode timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_random32_${N}gpu_$mode.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_random32_${N}gpu_$mode.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$mode", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "swaps=%s"%d.get("global_swaps_per_circuit"), "swap_bytes=%s"%r.get("swap_bytes_sent_per_rank"))
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/bench_random32_${N}gpu_$mode.log").read()[-1500:])
PY
done
