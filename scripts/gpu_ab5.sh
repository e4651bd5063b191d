#!/bin/bash
# A/B: one-tile-per-CTA vs persistent prefetch form of k_tile_pass, lazy |0..0> input on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
run() {
  local name=$1; shift
  for w in ${WORKLOADS:-qft30 hea28 random32}; do
    env "$@" timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_${name}_$w.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${name}_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$name", "$w", "gates/s=%.0f"%d["value"], "ms/step=%.2f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "hbm_pass_frac=%.3f"%r.get("hbm_pass_frac",0), d["clocks"]["reasons"])
except Exception as e:
    print("$name $w failed", e); print(open("gpurun_out/ab_${name}_$w.log").read()[-600:])
PY
  done
}
run base DVD_PERSIST=0 DVD_LAZY_ZERO=0
run lazy DVD_PERSIST=0 DVD_LAZY_ZERO=1
run persist DVD_PERSIST=1 DVD_LAZY_ZERO=0
run persist_lazy DVD_PERSIST=1 DVD_LAZY_ZERO=1
