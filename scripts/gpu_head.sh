#!/bin/bash
# round-1 HEAD confirmation: GPU tests, default bench (with CPU baseline), other configs, ncu launch list,
# DRAM traffic and a full capture of k_tile_pass on the 30-qubit QFT-style circuit.
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
T0=$SECONDS
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $? ($((SECONDS-T0)) s)"
tail -1 gpurun_out/bench_default.log | cut -c1-600
for w in hea28 random32 layered20; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.log").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$w", "gates/s=%.0f"%d["value"], "ms/step=%.1f"%d["ms_per_step"], "passes=%s"%d.get("passes_per_circuit"), "avg_launch_ms=%.2f"%r.get("avg_launch_ms",0), "hbm_pass_frac=%.3f"%r.get("hbm_pass_frac",0), "e2e=%.0f"%((d.get("e2e") or {}).get("value",0)))
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.log").read()[-800:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_qft30.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -c 3 -f -o gpurun_out/prof_tile_qft30 \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-400
