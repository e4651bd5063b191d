#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "jit_specialised and hea" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
r = d["roofline"]
print("value %.0f gates/s  %.2f ms/step | dense %.0f (%.2f ms) | e2e %.0f | hbm_pass_frac %.3f | avg_launch %.2f ms" % (d["value"], d["ms_per_step"], d["dense_state"]["value"], d["dense_state"]["ms_per_step"], d["e2e"]["value"], r["hbm_pass_frac"], r["avg_launch_ms"]))
print("jit", d["config"]["jit"]); print("scaling_point", (d.get("scaling_point") or {}).get("value")); print("cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d["gpu_launches"], d["clocks"])
for k, v in r["single_gate_pass"].items(): print("  ", k, "%.2f ms frac %.3f" % (v["ms"], v["hbm_frac"]))
PY
