#!/bin/bash
# N-GPU validation at HEAD: parity vs oracle on N ranks, cfg 4 (random32, strong scaling) through bench.py as the
# driver launches it, and cfg 5's circuit (hea34) when it fits.
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_N=${DIST_CHECK_N:-16,22} timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
grep -E "^n=|DIST_CHECK|Error|error" gpurun_out/dist_check_${N}gpu.log | tail -12
show() {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f, "gates/s=%.0f" % d["value"], "ms/step=%.1f" % d["ms_per_step"], "passes=%s" % d.get("passes_per_circuit"),
          "swaps=%s" % d.get("global_swaps_per_circuit"), "swap_bytes=%s" % r.get("swap_bytes_sent_per_rank"),
          "e2e=%s" % ((d.get("e2e") or {}).get("value")), d["clocks"]["reasons"])
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-1500:])
PY
}
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_random32_${N}gpu.log 2>&1
show gpurun_out/bench_random32_${N}gpu.log
DVD_SWAP=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_random32_${N}gpu_nccl.log 2>&1
show gpurun_out/bench_random32_${N}gpu_nccl.log
if [ "$N" -ge 4 ]; then
  timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload hea34 --steps 1 --warmup 1 > gpurun_out/bench_hea34_${N}gpu.log 2>&1
  show gpurun_out/bench_hea34_${N}gpu.log
fi
timeout 300 $TR --master-port 29515 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_reference_${N}gpu.log 2>&1
tail -1 gpurun_out/bench_reference_${N}gpu.log | cut -c1-300
