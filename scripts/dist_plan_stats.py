"""Swap schedule of the distributed planner for a workload (host only, no GPU)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damavand_b200 import _lib, circuits
from oracle.oracle import OracleCircuit
from tests.helpers import gate_array


def stats(name, world):
    n, build = circuits.workload(name)
    g = world.bit_length() - 1
    o = OracleCircuit.__new__(OracleCircuit); o.num_qubits = n; o.gates = []; o.observables = []
    build(o)
    arr, ng = gate_array(o)
    L = _lib.load()
    cap = 64 + 8 * ng + 4096
    out = (ctypes.c_int32 * cap)()
    perm = (ctypes.c_int32 * n)(*range(n))
    k = L.dvd_plan_distributed_debug(n, n - g, arr, ng, perm, 1, out, cap)
    assert k > 0, L.dvd_last_error()
    pos = 1; swaps = []; seg = []
    for _ in range(out[0]):
        kind, a, b, ngt = out[pos:pos + 4]; pos += 4
        if kind == 1:
            swaps.append((a, b))
        else:
            seg.append(ngt)
        pos += 3 * ngt
    print(f"{name} world={world}: gates {ng}, swaps {len(swaps)}, local segments {seg}")
    print("   swaps (global, local):", swaps[:40])


if __name__ == "__main__":
    name = sys.argv[1]
    for w in [int(x) for x in sys.argv[2:]] or [8]:
        stats(name, w)
