#!/bin/bash
# round 2, session D (2 GPUs): distributed parity in every swap mode (pytest), then bench.py --gpus 2 (random32)
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2d_gpus.log 2>&1
N=${1:-2}
timeout 1500 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2d_pytest_dist.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2d_pytest_dist.log; tail -25 gpurun_out/r2d_pytest_dist.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2d_bench_random32_${N}gpu.log 2>&1
echo "bench exit $? ($((SECONDS-T0)) s)"
python - gpurun_out/r2d_bench_random32_${N}gpu.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("gates/s=%.0f ms/step=%.2f from_reset=%.2f passes=%s swaps=%s frac=%.3f" % (d["value"], d["ms_per_step"], d["from_reset"]["ms_per_step"], d.get("passes_per_circuit"), d.get("global_swaps_per_circuit"), r.get("frac", 0)))
    print("parity", json.dumps(d.get("parity"))[:1200])
    print("nvlink", json.dumps(r.get("nvlink")))
    print("strong_scaling", json.dumps(d.get("strong_scaling"))[:1500])
    print("e2e", d.get("e2e"), "sanity", d.get("sanity"))
    print("jit", (d["config"].get("jit") or {}).get("final"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-3000:])
PY
echo "total $((SECONDS-T0)) s"
