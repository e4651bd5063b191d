#!/bin/bash
mkdir -p gpurun_out
run() {
  local name=$1; shift
  for w in ${WORKLOADS:-qft30 hea28}; do
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
run s0 DVD_STAGGER=0
run s05 DVD_STAGGER=0.5
run s10 DVD_STAGGER=1.0
run s20 DVD_STAGGER=2.0
python - <<'PY' > gpurun_out/e2e_breakdown.log 2>&1
import time, numpy as np, sys
sys.path.insert(0, ".")
from damavand_b200 import Circuit, circuits
for name in ["qft30", "random32"]:
    n, build = circuits.workload(name)
    c = Circuit(n, "gpu"); build(c)
    for q in range(n): c.add_pauli_z_gate(q, True)
    u = np.random.default_rng(1).random(1000)
    for it in range(3):
        t0 = time.perf_counter(); c.reset_amplitudes(); c.synchronize()
        t1 = time.perf_counter(); c.forward()
        t2 = time.perf_counter(); s = c.sample_numpy(1000, u)
        t3 = time.perf_counter(); ev = c.extract_expectation_values_numpy(s)
        t4 = time.perf_counter()
        print(name, it, "reset %.1f forward %.1f sample %.1f extract %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), flush=True)
    c.close()
PY
cat gpurun_out/e2e_breakdown.log
