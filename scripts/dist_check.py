"""Multi-GPU parity check of apply_method="distributed_gpu" against the CPU oracle.

Launch (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py
Every rank compares ITS chunk with the oracle's slice (amplitudes within 1e-12 relative), the
allreduced norm and <Z_q>, and the distributed sampler's indices bit for bit with the restatement of
sample_distributed (circuit_distributed.rs:42-129) fed the same uniforms.  Test infrastructure: uses
oracle/.  Prints one line per case on rank 0 and exits non-zero on any mismatch.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    # every rank runs the CPU oracle itself: share the host cores instead of oversubscribing them
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world_env))))
    from damavand_b200 import Circuit, circuits, distributed
    from oracle import oracle
    from oracle.oracle import OracleCircuit
    rank, world = distributed.initialize("nccl")
    import torch.distributed as dist
    fails = 0
    cases = []
    for n in (int(x) for x in os.environ.get("DIST_CHECK_N", "14,17,21,24").split(",")):
        cases += [(n, "random", 300), (n, "hea", 4)]
        if n <= int(os.environ.get("DIST_CHECK_QFT_MAX", "21")):        # (n^2 / 2 gates: the CPU oracle's time, not the GPU's)
            cases += [(n, "qft", 0), (n, "layered", 3)]
    jit = int(os.environ.get("DIST_CHECK_JIT", "0"))
    for n, kind, arg in cases:
        g = Circuit(n, "distributed_gpu")
        if jit:
            g.set_jit(2)          # run-time specialised kernels, compiled on first use
        o = OracleCircuit(n)
        for c in (g, o):
            if kind == "random":
                circuits.random_circuit(c, n, arg, seed=n)
            elif kind == "qft":
                circuits.qft_like(c, n)
            elif kind == "hea":
                circuits.hea(c, n, arg)
            else:
                circuits.layered(c, n, arg)
            for q in range(n):
                c.add_pauli_z_gate(q, True)
        g.forward()
        o.forward()
        chunk = (1 << n) // world
        want = o.amplitudes()[rank * chunk:(rank + 1) * chunk]
        got = g.state_numpy()
        err = float(np.abs(got - want).max() / np.abs(o.amplitudes()).max())
        l2 = float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300))
        p = o.measure_np()
        idx = np.arange(p.size)
        ez_want = np.array([(p * (1 - 2.0 * ((idx >> q) & 1))).sum() for q in range(n)])
        ez_err = float(np.abs(g.expectation_z() - ez_want).max())
        norm_err = abs(g.norm() - 1.0)
        shots = 5000
        rng = np.random.default_rng(1235)
        u = rng.random(2 * shots)
        s = g.sample_numpy(shots, u)
        s_want = oracle.sample_distributed(p, world, u[:shots], u[shots:], "tree")
        s_ok = bool((s == s_want).all())
        ev_ok = g.extract_expectation_values(s[:50].tolist()) == o.extract_expectation_values(s[:50].tolist())
        st = g.stats()
        # a second forward starts from a dense state in the layout the first one restored
        g.forward(); o.forward()
        got2 = g.state_numpy()
        err2 = float(np.abs(got2 - o.amplitudes()[rank * chunk:(rank + 1) * chunk]).max() / np.abs(o.amplitudes()).max())
        st = g.stats()
        ok = err < 1e-12 and l2 < 1e-12 and err2 < 1e-12 and ez_err < 1e-12 and norm_err < 1e-12 and s_ok and ev_ok
        import torch
        t = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            print(f"n={n} world={world} {kind:8s} err={err:.2e} l2={l2:.2e} err_2nd_forward={err2:.2e} ez_err={ez_err:.2e} norm_err={norm_err:.1e} "
                  f"samples_ok={s_ok} ev_ok={ev_ok} swaps={st['global_swaps']} fused_remap_passes={st['remap_passes']} store_side={st['store_remap_passes']} "
                  f"jit_launches={st['jit_launches']} passes={st['tile_passes']} "
                  f"simple={st['simple_passes']} all_ranks_ok={int(t.item()) == 0}", flush=True)
        fails += int(t.item())
        g.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if fails == 0 else f"FAIL ({fails})", flush=True)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
