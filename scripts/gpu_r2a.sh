#!/bin/bash
# round 2, session A (1 GPU): parity of the new kernel forms, then A/B of the forms on the three single-GPU workloads
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "kernel_forms or jit" > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)" >> gpurun_out/r2a_pytest.log; tail -15 gpurun_out/r2a_pytest.log | cut -c1-400
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f" % d["value"], "ms/step=%.2f" % d["ms_per_step"], "from_reset=%.2f ms" % d["from_reset"]["ms_per_step"],
          "passes=%s" % d.get("passes_per_circuit"), "avg_launch_ms=%.3f" % r.get("avg_launch_ms", 0), "hbm_pass_frac=%.3f" % r.get("hbm_pass_frac", 0),
          "jit=%s" % d["config"].get("jit"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-1500:])
PY
}
for w in qft30 hea28 random32; do
  forms="classic2 classic3 ring auto"
  [ $w = hea28 ] && forms="classic2 ring auto"
  [ $w = random32 ] && forms="classic2 ring"
  for form in $forms; do
    if [ $form = auto ]; then unset DVD_JIT_FORM; else export DVD_JIT_FORM=$form; fi
    timeout 300 python bench.py --workload $w --jit 1 --steps 3 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate > gpurun_out/r2a_${w}_$form.log 2>&1
    show gpurun_out/r2a_${w}_$form.log
  done
done
unset DVD_JIT_FORM
# interpreter kernels: one tile per CTA against the ring form
for w in qft30; do
  for ring in 0 1; do
    DVD_RING=$ring timeout 300 python bench.py --workload $w --jit 0 --steps 3 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate > gpurun_out/r2a_${w}_interp_ring$ring.log 2>&1
    show gpurun_out/r2a_${w}_interp_ring$ring.log
  done
done
echo "total $((SECONDS-T0)) s"
