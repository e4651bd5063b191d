#!/bin/bash
# round 2, session I (2 GPUs): what the NVLink side of a pass costs in each form -- fused remap with the classic and the ring
# kernel form, and the stand-alone in-place exchange (k_swap_peer)
mkdir -p gpurun_out
N=2
run() {
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 2 --no-parity --no-scaling-point --no-e2e > gpurun_out/r2i_$tag.log 2>&1
  python - gpurun_out/r2i_$tag.log <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    nv = r.get("nvlink") or {}
    print(f.split("/")[-1], "ms/step=%.2f from_reset=%.2f passes=%s swaps=%s" % (d["ms_per_step"], d["from_reset"]["ms_per_step"], d.get("passes_per_circuit"), d.get("global_swaps_per_circuit")),
          "nvlink ms=%.2f GB/s/dir=%.0f fused=%s avg_fused=%s avg_plain=%s" % (nv.get("ms") or 0, nv.get("gbs_per_dir") or 0, nv.get("fused_remap_passes"), nv.get("avg_fused_pass_ms"), nv.get("avg_plain_pass_ms")),
          "chosen=%s" % ((d["config"].get("jit") or {}).get("final") or {}).get("chosen"))
except Exception as e:
    print(f, "failed", e); print(open(f).read()[-2000:])
PY
}
run fused_classic2 DVD_JIT_FORM=classic2
run fused_ring DVD_JIT_FORM=ring
run inplace_swap DVD_FUSED_REMAP=0 DVD_JIT_FORM=classic2
