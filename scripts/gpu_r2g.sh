#!/bin/bash
# round 2, session G (1 GPU): GPU suite with tile relabelling on by default (26-qubit cases excluded here), the pure-I/O
# pass in every kernel form, the default bench line, and the ncu launch list of a short bench run
mkdir -p gpurun_out
T0=$SECONDS
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "not config_scale" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2g_pytest.log; tail -8 gpurun_out/r2g_pytest.log | cut -c1-400
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.3f" % d["ms_per_step"], "from_reset=%.2f ms" % d["from_reset"]["ms_per_step"],
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.3f" % r.get("avg_launch_ms", 0), "frac=%.3f" % r.get("frac", 0),
          "fp64=%.3f" % ((r.get("fp64") or {}).get("frac") or 0), "e2e=%s" % ((d.get("e2e") or {}).get("value")),
          "jit=%s" % ((d["config"].get("jit") or {}).get("final") or {}).get("chosen"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-2500:])
PY
}
for form in classic2 classic3 ring; do
  DVD_LAZY_ZERO=0 DVD_JIT_FORM=$form timeout 300 python bench.py --workload hhi30 --steps 5 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2g_hhi30_$form.log 2>&1
  show gpurun_out/r2g_hhi30_$form.log
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_default.log 2>&1
show gpurun_out/r2g_bench_default.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_qft30.csv \
    python bench.py --steps 2 --warmup 1 --no-scaling-point --no-cpu-baseline --no-parity > gpurun_out/r2g_ncu_launches.log 2>&1
tail -2 gpurun_out/r2g_ncu_launches.log | cut -c1-300
echo "total $((SECONDS-T0)) s"
