#!/bin/bash
# round 2, session R (4 GPUs): the final schedule (tail deferral, push on dense states, pull while qubits are still |0>)
# on four ranks: parity leg on all ranks, dense forward, from a reset, strong scaling against the 1-GPU base.
mkdir -p gpurun_out
N=${1:-4}
T0=$SECONDS
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/r2r_bench_${N}gpu.log 2>&1
echo "bench exit $? ($((SECONDS-T0)) s)"
python - gpurun_out/r2r_bench_${N}gpu.log <<'PY'
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
echo "total $((SECONDS-T0)) s"
