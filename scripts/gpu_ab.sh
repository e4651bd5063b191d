#!/bin/bash
# A/B run on one B200: GPU parity tests on the current build, then the three single-GPU workloads under
# several build / option variants.  Lines: <variant> <workload> gates/s ms/step passes avg_launch_ms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
run() {  # name, env...
  local name=$1; shift
  for w in ${WORKLOADS:-qft30 hea28 random32}; do
    env "$@" timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ab_${name}_$w.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${name}_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$name", "$w", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "hbm_pass_frac=%.3f"%r.get("hbm_pass_frac",0))
except Exception as e:
    print("$name $w failed", e); print(open("gpurun_out/ab_${name}_$w.log").read()[-600:])
PY
  done
}
run v4 DVD_LIB_PATH=$PWD/damavand_b200/libdvd_v4.so
run v5c1 DVD_PLAN_CANDIDATES=1
run v5 DVD_PLAN_CANDIDATES=12
run v5s3 DVD_STAGGER=0.3
run v5s5 DVD_STAGGER=0.5
