#!/bin/bash
# round 2, session N (1 GPU): do the CTAs that share an SM run in lockstep?  Stagger the first waves by a fraction of a tile time.
mkdir -p gpurun_out
run() {
  w=$1; st=$2
  DVD_STAGGER=$st DVD_JIT_FORM=classic2 timeout 300 python bench.py --workload $w --steps 5 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2n_${w}_st$st.log 2>&1
  python - gpurun_out/r2n_${w}_st$st.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    print(f.split("/")[-1], "ms/step=%.3f" % d["ms_per_step"], "from_reset=%.3f" % d["from_reset"]["ms_per_step"], "frac=%.3f" % d["roofline"]["frac"])
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-1500:])
PY
}
for st in 0 4000 9000 14000; do run qft30 $st; done
for st in 0 16000; do run hea28 $st; done
for st in 0 12000; do run random32 $st; done
