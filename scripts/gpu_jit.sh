#!/bin/bash
# structure-specialised (NVRTC) pass kernels: parity test, then the three single-GPU workloads with --jit 1
mkdir -p gpurun_out
T0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k jit_specialised > gpurun_out/pytest_jit.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/pytest_jit.log; tail -25 gpurun_out/pytest_jit.log | cut -c1-300
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.2f" % d["ms_per_step"], "dense=%.0f (%.2f ms)" % (d["dense_state"]["value"], d["dense_state"]["ms_per_step"]),
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.2f" % r.get("avg_launch_ms", 0), "hbm_pass_frac=%.3f" % r.get("hbm_pass_frac", 0),
          "e2e=%.0f" % ((d.get("e2e") or {}).get("value") or 0), "jit=%s" % d["config"].get("jit"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-1500:])
PY
}
for w in qft30 hea28 random32; do
  timeout 300 python bench.py --workload $w --jit 1 --steps 3 --warmup 2 --no-cpu-baseline --no-scaling-point > gpurun_out/bench_jit_$w.log 2>&1
  show gpurun_out/bench_jit_$w.log
done
