#!/bin/bash
# round 2, session Q (2 GPUs): push against pull -- every swap round on the STORE of the pass in front of it (DVD_STORE_REMAP=2)
# against the default (only the layout restore rides on a store); the store-side remap as a compile-time kernel variant.
mkdir -p gpurun_out
N=${1:-2}
T0=$SECONDS
env DVD_STORE_REMAP=2 DIST_CHECK_JIT=1 DIST_CHECK_N=14,18,20 DIST_CHECK_QFT_MAX=20 DVD_JIT_MIN_QUBITS=12 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29531 scripts/dist_check.py > gpurun_out/r2q_dist_check_push.log 2>&1
echo "dist_check push exit $? ($((SECONDS-T0)) s): $(grep -c all_ranks_ok=True gpurun_out/r2q_dist_check_push.log) ok lines, store-side in $(grep -c 'store_side=[1-9]' gpurun_out/r2q_dist_check_push.log); $(grep DIST_CHECK gpurun_out/r2q_dist_check_push.log)"
grep -v all_ranks_ok=True gpurun_out/r2q_dist_check_push.log | grep '^n=' | head -5
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print(f.split("/")[-1], "gates/s=%.0f ms/step=%.2f from_reset=%.2f passes=%s swaps=%s frac=%.3f" % (d["value"], d["ms_per_step"], d["from_reset"]["ms_per_step"], d.get("passes_per_circuit"), d.get("global_swaps_per_circuit"), r.get("frac", 0)))
    print("  parity ok", (d.get("parity") or {}).get("ok"), [(c["n"], c["circuit"], "%.1e" % c["max_rel_err"], c["samples_ok"], c["fused_remap_passes"], c.get("store_side_remap_passes"), c["passes"]) for c in (d.get("parity") or {}).get("cases", [])])
    nv = r.get("nvlink") or {}
    print("  nvlink", {k: nv.get(k) for k in ("bytes_per_dir", "ms", "gbs_per_dir", "fused_remap_passes", "store_side_remap_passes", "avg_store_side_pass_ms", "avg_load_side_pass_ms", "avg_plain_pass_ms")})
    ss = d.get("strong_scaling") or {}
    print("  strong_scaling eff", ss.get("efficiency"), "base", (ss.get("base") or {}).get("value"), (ss.get("base") or {}).get("ms_per_step"), "sanity", d.get("sanity"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-3000:])
PY
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/r2q_bench_${N}gpu_default.log 2>&1
echo "bench default exit $? ($((SECONDS-T0)) s)"; show gpurun_out/r2q_bench_${N}gpu_default.log
DVD_STORE_REMAP=2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-scaling-point > gpurun_out/r2q_bench_${N}gpu_push.log 2>&1
echo "bench push exit $? ($((SECONDS-T0)) s)"; show gpurun_out/r2q_bench_${N}gpu_push.log
echo "total $((SECONDS-T0)) s"
