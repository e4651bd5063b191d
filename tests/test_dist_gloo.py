"""N>1 path on CPU: world_size-2 (and 4) gloo process groups execute the REAL distributed schedule
(dvd_plan_distributed_debug = the planner the CUDA engine runs) on numpy shards, with the half-chunk
exchanges done by torch.distributed send/recv, and compare with the single-process oracle.
Also covers the distributed sampling protocol (per-rank totals -> rank draw -> local draw -> reduce)."""
import ctypes
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from damavand_b200 import _lib, circuits, distributed
from oracle import oracle
from oracle.oracle import OracleCircuit
from tests.helpers import gate_array


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _half_index(h, lq, bit):
    return ((h >> lq) << (lq + 1)) | (bit << lq) | (h & ((1 << lq) - 1))


def _apply_local(shard, n_local, rank, target, control, m):
    """One physical-qubit gate on this rank's shard (interleaved float64), reference update rule."""
    n = 1 << n_local
    L = oracle.lib()
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if control >= n_local:                      # rank-index control: a per-rank predicate
        if not (rank >> (control - n_local)) & 1:
            return
        control = -1
    if target >= n_local:                       # rank-index target: diagonal gates only
        assert m[2] == m[3] == m[4] == m[5] == 0.0
        bit = (rank >> (target - n_local)) & 1
        d = complex(m[6], m[7]) if bit else complex(m[0], m[1])
        z = shard.view(np.complex128)
        if control < 0:
            z *= d
        else:
            idx = np.arange(n)
            z[((idx >> control) & 1) == 1] *= d
        return
    scratch = np.empty_like(shard)
    mm = np.ascontiguousarray(np.array(m, dtype=np.float64))
    L.orc_apply_gate(dp(shard), dp(scratch), n, dp(mm), int(control), int(target))


def _worker(rank, world, port, n, workload, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w = distributed.initialize("gloo")
    assert (r, w) == (rank, world)
    try:
        g = world.bit_length() - 1
        n_local = n - g
        chunk = 1 << n_local
        circ = OracleCircuit(n)
        if workload == "random":
            circuits.random_circuit(circ, n, 150, seed)
        elif workload == "qft":
            circuits.qft_like(circ, n)
        else:
            circuits.hea(circ, n, 3, seed)
        arr, ng = gate_array(circ)
        L = _lib.load()
        perm = (ctypes.c_int32 * n)(*range(n))
        cap = 64 + 8 * ng + 4096
        out = (ctypes.c_int32 * cap)()
        k = L.dvd_plan_distributed_debug(n, n_local, arr, ng, perm, 1, out, cap)
        assert k > 0 and list(perm) == list(range(n))
        shard = np.zeros(2 * chunk, dtype=np.float64)
        if rank == 0:
            shard[0] = 1.0
        pos, n_swaps = 1, 0
        for _ in range(out[0]):
            kind, a, b, cnt = out[pos], out[pos + 1], out[pos + 2], out[pos + 3]
            pos += 4
            if kind == 1:      # GLOBAL_SWAP(gq=a, lq=b)
                n_swaps += 1
                j = a - n_local
                bit = (rank >> j) & 1
                partner = rank ^ (1 << j)
                assert partner == oracle.lib().orc_compute_partner_rank(rank, chunk, chunk << j)   # circuit.rs:781-795
                h = np.arange(chunk // 2)
                idx = _half_index(h, b, 1 - bit)
                z = shard.view(np.complex128)
                send = torch.from_numpy(np.ascontiguousarray(z[idx]).view(np.float64).copy())
                recv = torch.empty_like(send)
                ops = [dist.P2POp(dist.isend, send, partner), dist.P2POp(dist.irecv, recv, partner)]
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                z[idx] = recv.numpy().view(np.complex128)
            else:
                for _g in range(cnt):
                    gi, t, c = out[pos], out[pos + 1], out[pos + 2]
                    pos += 3
                    m = list(arr[gi].m) if gi >= 0 else [0, 0, 1, 0, 1, 0, 0, 0]   # layout-restoring CNOT
                    _apply_local(shard, n_local, rank, t, c, m)
        # gather on rank 0 and compare with the single-process oracle
        full = [torch.empty(2 * chunk, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(shard), full, dst=0)
        # distributed sampling protocol on the shards (tree order), mirrors dvd_sample for world > 1
        shots = 4000
        rng = np.random.default_rng(77)
        u_rank, u_loc = rng.random(shots), rng.random(shots)
        z = shard.view(np.complex128)
        p_loc = z.real * z.real + z.imag * z.imag
        tot = torch.tensor([oracle.tree_total(p_loc)], dtype=torch.float64)
        tots = [torch.empty(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(tots, tot)
        totals = np.array([float(t) for t in tots])
        sel = oracle.sample_sequential(totals, u_rank)
        mine = np.nonzero(sel == rank)[0]
        res = torch.zeros(shots, dtype=torch.int64)
        if mine.size:
            res[mine] = torch.from_numpy((oracle.sample_tree(p_loc, u_loc[mine]) + np.uint64(rank * chunk)).astype(np.int64))
        dist.all_reduce(res)
        if rank == 0:
            circ.forward()
            got = torch.cat(full).numpy().view(np.complex128)
            err = float(np.abs(got - circ.amplitudes()).max() / np.abs(circ.amplitudes()).max())
            want = oracle.sample_distributed(circ.measure_np(), world, u_rank, u_loc, "tree")
            q.put((err, n_swaps, bool((res.numpy().astype(np.uint64) == want).all())))
    finally:
        dist.barrier()
        dist.destroy_process_group()


# (n_local >= 12: the schedule comes from the tuned planner, tail deferral included)
@pytest.mark.parametrize("world,n,workload", [(2, 12, "random"), (2, 13, "qft"), (4, 12, "random"), (2, 12, "hea"),
                                              (2, 14, "random"), (4, 15, "hea")])
def test_distributed_schedule_over_gloo(world, n, workload):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, workload, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    for p in procs:
        assert p.exitcode == 0
    err, n_swaps, samples_ok = q.get(timeout=5)
    assert err < 1e-12
    assert n_swaps >= 1
    assert samples_ok
