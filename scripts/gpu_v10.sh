#!/bin/bash
# support-tracking build: GPU tests, default bench (with CPU baseline + scaling point), other configs
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
T0=$SECONDS
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $? ($((SECONDS-T0)) s)"
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.2f" % d["ms_per_step"], "dense=%.0f (%.2f ms)" % (d["dense_state"]["value"], d["dense_state"]["ms_per_step"]),
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.2f" % r.get("avg_launch_ms", 0), "hbm_pass_frac=%.3f" % r.get("hbm_pass_frac", 0),
          "e2e=%.0f" % ((d.get("e2e") or {}).get("value") or 0), d["clocks"]["reasons"])
    for k, v in (r.get("single_gate_pass") or {}).items(): print("   ", k, "%.2f ms  %.0f GB/s  frac %.3f" % (v["ms"], v["hbm_gbs"], v["hbm_frac"]))
    if d.get("scaling_point"): print("    scaling_point", d["scaling_point"].get("value"), d["scaling_point"].get("ms_per_step"))
    if d.get("cpu_baseline"): print("    cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-1200:])
PY
}
show gpurun_out/bench_default.log
for w in hea28 random32; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  show gpurun_out/bench_$w.log
done
