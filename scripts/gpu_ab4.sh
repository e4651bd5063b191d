#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
run() {
  local name=$1; shift
  for w in ${WORKLOADS:-qft30 hea28 random32}; do
    env "$@" timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ab_${name}_$w.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${name}_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$name", "$w", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "hbm_pass_frac=%.3f"%r.get("hbm_pass_frac",0), d["clocks"]["reasons"])
except Exception as e:
    print("$name $w failed", e); print(open("gpurun_out/ab_${name}_$w.log").read()[-600:])
PY
  done
}
run pf0 DVD_LIB_PATH=$PWD/damavand_b200/libdvd_pf0.so
run pf1 DVD_MACRO_OPS=1
