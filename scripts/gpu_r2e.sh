#!/bin/bash
# round 2, session E (1 GPU): new parity tests (26 qubits, sampler orders, unitary gates, profiling, fidelity, <Z> beyond 32
# local qubits), the pure-I/O pass at 2 / 3 CTAs per SM / ring, hea28 with and without tile relabelling
mkdir -p gpurun_out
T0=$SECONDS
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "config_scale or unitary or profiling or fidelity or beyond_32 or jit_specialised or load_state or read" > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2e_pytest.log; tail -15 gpurun_out/r2e_pytest.log | cut -c1-400
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.3f" % d["ms_per_step"], "from_reset=%.2f ms" % d["from_reset"]["ms_per_step"],
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.3f" % r.get("avg_launch_ms", 0), "frac=%.3f" % r.get("frac", 0),
          "jit=%s" % ((d["config"].get("jit") or {}).get("final") or {}).get("chosen"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-2500:])
PY
}
for form in classic2 classic3 ring; do
  DVD_JIT_FORM=$form timeout 300 python bench.py --workload hhi30 --steps 5 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2e_hhi30_$form.log 2>&1
  show gpurun_out/r2e_hhi30_$form.log
done
for rl in 0 1; do
  DVD_RELABEL=$rl timeout 400 python bench.py --workload hea28 --steps 3 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2e_hea28_relabel$rl.log 2>&1
  show gpurun_out/r2e_hea28_relabel$rl.log
done
echo "total $((SECONDS-T0)) s"
