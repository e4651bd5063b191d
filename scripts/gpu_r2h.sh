mkdir -p gpurun_out
for w in hhi30x4 hhi30x8 hhi30x9; do
  DVD_LAZY_ZERO=0 DVD_JIT_FORM=classic2 timeout 300 python bench.py --workload $w --steps 5 --warmup 2 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate --no-parity > gpurun_out/r2h_$w.log 2>&1
  python - gpurun_out/r2h_$w.log <<'PY'
import json, sys
f = sys.argv[1]
d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
print(f.split("/")[-1], "ms/step=%.3f" % d["ms_per_step"], "frac=%.3f" % d["roofline"]["frac"])
PY
done
