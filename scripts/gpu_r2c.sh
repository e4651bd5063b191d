#!/bin/bash
# round 2, session C (1 GPU): whole GPU test suite, then the three single-GPU workloads with the offset stash
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2c_pytest.log; tail -15 gpurun_out/r2c_pytest.log | cut -c1-400
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.2f" % d["ms_per_step"], "from_reset=%.2f ms" % d["from_reset"]["ms_per_step"],
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.3f" % r.get("avg_launch_ms", 0), "frac=%.3f" % r.get("frac", 0),
          "parity=%s" % ((d.get("parity") or {}).get("ok")), "sanity=%s" % d.get("sanity"), "jit=%s" % (d["config"].get("jit") or {}).get("final"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-2500:])
PY
}
for w in qft30 hea28 random32; do
  for form in classic2 auto; do
    if [ $form = auto ]; then unset DVD_JIT_FORM; else export DVD_JIT_FORM=$form; fi
    extra="--no-parity"; [ $form = auto ] && [ $w = qft30 ] && extra=""
    timeout 400 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate $extra > gpurun_out/r2c_${w}_$form.log 2>&1
    show gpurun_out/r2c_${w}_$form.log
  done
done
echo "total $((SECONDS-T0)) s"
