#!/bin/bash
# iteration run: GPU tests + the three single-GPU benches (+ optional ncu capture when NCU=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
for w in qft30 hea28 random32; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$w", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "hbm_frac=%.3f"%r.get("frac",0), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "e2e=%.0f"%((d.get("e2e") or {}).get("value",0)))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.log").read()[-800:])
PY
done
if [ "$NCU" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 3 -c 3 -f -o gpurun_out/prof_tile_qft26 \
   python bench.py --workload qft26 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
fi
