#!/bin/bash
# final round-1 confirmation: JIT parity tests, large-state / pipeline tests with every pass on a JIT kernel, default bench
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k jit_specialised > gpurun_out/pytest_jit.log 2>&1
echo "pytest jit exit $? ($((SECONDS-T0)) s)" >> gpurun_out/pytest_jit.log; tail -4 gpurun_out/pytest_jit.log | cut -c1-300
T0=$SECONDS
DVD_JIT=sync DVD_JIT_MIN_QUBITS=12 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "large_state or shapes_qft_and_hea or cfg1 or fused_equals" > gpurun_out/pytest_jit_all.log 2>&1
echo "pytest jit-everywhere exit $? ($((SECONDS-T0)) s)" >> gpurun_out/pytest_jit_all.log; tail -4 gpurun_out/pytest_jit_all.log | cut -c1-300
T0=$SECONDS
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $? ($((SECONDS-T0)) s)"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
r = d["roofline"]
print("value %.0f gates/s  %.2f ms/step | dense %.0f (%.2f ms) | e2e %.0f | hbm_pass_frac %.3f | avg_launch %.2f ms" % (d["value"], d["ms_per_step"], d["dense_state"]["value"], d["dense_state"]["ms_per_step"], d["e2e"]["value"], r["hbm_pass_frac"], r["avg_launch_ms"]))
print("jit", d["config"]["jit"]); print("scaling_point", d.get("scaling_point", {}).get("value")); print("cpu", (d.get("cpu_baseline") or {}).get("value"))
for k, v in r["single_gate_pass"].items(): print("  ", k, "%.2f ms frac %.3f" % (v["ms"], v["hbm_frac"]))
PY
