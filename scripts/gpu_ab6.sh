#!/bin/bash
# A/B of kernel builds: base (store base reuse) / Hadamard fma form / + table constants before the tile loads
mkdir -p gpurun_out
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
run base DVD_LAZY_ZERO=1
run hadfma DVD_LIB_PATH=$PWD/damavand_b200/libdvd_hadfma.so
run hadfma_wc DVD_LIB_PATH=$PWD/damavand_b200/libdvd_hadfma_wc.so
WORKLOADS="qft30" run hadfma_wc_persist DVD_LIB_PATH=$PWD/damavand_b200/libdvd_hadfma_wc.so DVD_PERSIST=1
DVD_LIB_PATH=$PWD/damavand_b200/libdvd_hadfma_wc.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "random_circuits or shapes or kernel_forms or large_state" 2>&1 | tail -3
