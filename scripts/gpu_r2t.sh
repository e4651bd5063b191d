#!/bin/bash
# round 2, session T (1 GPU, the last GPU-minute): the headline workload at HEAD without the side legs
mkdir -p gpurun_out
timeout 58 python bench.py --workload ${1:-qft30} --steps ${2:-10} --warmup 3 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2t_bench_${1:-qft30}_lean.log 2>&1
echo "bench exit $? ($SECONDS s)"
python - <<'PY'
import json
import sys, glob
f = sorted(glob.glob("gpurun_out/r2t_bench_*_lean.log"), key=__import__("os").path.getmtime)[-1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("gates/s=%.0f ms/step=%.3f from_reset=%.3f passes=%s frac=%.3f kernel=%s clocks=%s sanity=%s" % (d["value"], d["ms_per_step"], d["from_reset"]["ms_per_step"], d.get("passes_per_circuit"), r.get("frac", 0), r.get("kernel"), d.get("clocks"), d.get("sanity")))
except Exception as e:
    print("failed", e); print(open(f).read()[-2000:])
PY
