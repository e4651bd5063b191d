#!/bin/bash
# round 2, session P (2 GPUs): tail deferral + store-side layout restore on hardware: oracle parity on both ranks
# (interpreter and run-time specialised kernels, plain schedule as the control), then the bench line.
mkdir -p gpurun_out
N=${1:-2}
T0=$SECONDS
chk() {
  tag=$1; shift
  env "$@" DIST_CHECK_N=14,18,20 DIST_CHECK_QFT_MAX=20 DVD_JIT_MIN_QUBITS=12 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29531 scripts/dist_check.py > gpurun_out/r2p_dist_check_${tag}.log 2>&1
  echo "dist_check $tag exit $? ($((SECONDS-T0)) s): $(grep -c all_ranks_ok=True gpurun_out/r2p_dist_check_${tag}.log) ok lines, store-side in $(grep -c 'store_side=[1-9]' gpurun_out/r2p_dist_check_${tag}.log); $(grep DIST_CHECK gpurun_out/r2p_dist_check_${tag}.log)"
  grep -v all_ranks_ok=True gpurun_out/r2p_dist_check_${tag}.log | grep '^n=' | head -5
}
chk jit DIST_CHECK_JIT=1
chk interp DIST_CHECK_JIT=0

timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2p_bench_${N}gpu.log 2>&1
echo "bench exit $? ($((SECONDS-T0)) s)"
python - gpurun_out/r2p_bench_${N}gpu.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("gates/s=%.0f ms/step=%.2f from_reset=%.2f passes=%s swaps=%s frac=%.3f fp64=%.3f" % (d["value"], d["ms_per_step"], d["from_reset"]["ms_per_step"], d.get("passes_per_circuit"), d.get("global_swaps_per_circuit"), r.get("frac", 0), (r.get("fp64") or {}).get("frac", 0)))
    print("parity ok", (d.get("parity") or {}).get("ok"), [(c["n"], c["circuit"], c["max_rel_err"], c["samples_ok"], c["fused_remap_passes"], c.get("store_side_remap_passes"), c["passes"]) for c in (d.get("parity") or {}).get("cases", [])])
    print("nvlink", json.dumps(r.get("nvlink")))
    ss = d.get("strong_scaling") or {}
    print("strong_scaling eff", ss.get("efficiency"), "base", (ss.get("base") or {}).get("value"), (ss.get("base") or {}).get("ms_per_step"))
    print("cfg5", json.dumps(d.get("cfg5_point"))[:2500])
    print("e2e", d.get("e2e"), "sanity", d.get("sanity"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-3000:])
PY
echo "total $((SECONDS-T0)) s"
