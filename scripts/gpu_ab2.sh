#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for w in qft30 hea28 random32 layered20; do
    timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_v6_$w.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_v6_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("v6", "$w", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "hbm_pass_frac=%.3f"%r.get("hbm_pass_frac",0), "e2e=%.0f"%(d.get("e2e") or {}).get("value",0))
except Exception as e:
    print("v6 $w failed", e); print(open("gpurun_out/ab_v6_$w.log").read()[-600:])
PY
done
timeout 600 python scripts/opcost.py 28 > gpurun_out/opcost_v6.log 2>&1; cat gpurun_out/opcost_v6.log
