#!/bin/bash
# round 2, session L (1 GPU): the whole GPU suite at HEAD, the cfg 2 / cfg 4 single-GPU lines with both rooflines, the default line
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2l_pytest.log; tail -6 gpurun_out/r2l_pytest.log | cut -c1-400
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.3f" % d["ms_per_step"], "from_reset=%.2f ms" % d["from_reset"]["ms_per_step"],
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.3f" % r.get("avg_launch_ms", 0), "frac=%.3f" % r.get("frac", 0),
          "fp64=%s" % json.dumps(r.get("fp64")), "e2e=%s" % ((d.get("e2e") or {}).get("value")), "parity=%s" % ((d.get("parity") or {}).get("ok")))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-2500:])
PY
}
for w in hea28 random32; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-scaling-point --no-parity > gpurun_out/r2l_bench_$w.log 2>&1
  show gpurun_out/r2l_bench_$w.log
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_default.log 2>&1
show gpurun_out/r2l_bench_default.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_reference.log 2>&1
tail -1 gpurun_out/r2l_bench_reference.log | cut -c1-600
echo "total $((SECONDS-T0)) s"
