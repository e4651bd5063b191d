"""Print the planner's schedule statistics for a workload (host only, no GPU)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damavand_b200 import _lib, circuits
from tests.helpers import Recorder
from oracle.oracle import OracleCircuit

def stats(name, fuse=1, n_local=None):
    n, build = circuits.workload(name)
    o = OracleCircuit.__new__(OracleCircuit); o.num_qubits = n; o.gates = []; o.observables = []
    build(o)
    from tests.helpers import gate_array
    arr, ng = gate_array(o)
    L = _lib.load()
    cap = 64 + 7 * ng + 20 * 4096
    out = (ctypes.c_int32 * cap)()
    k = L.dvd_plan_debug(n, n_local or n, arr, ng, fuse, out, cap)
    assert k > 0, L.dvd_last_error()
    pos = 1; passes = []
    for _ in range(out[0]):
        tile = list(out[pos:pos+12]); sw = out[pos+12]; nops = out[pos+13]; pos += 14
        kinds = {}
        for i in range(nops):
            kd = out[pos+1]; kinds[kd] = kinds.get(kd, 0) + 1; pos += 5
        passes.append((tile, sw, nops, kinds))
    tot = sum(p[2] for p in passes)
    print(f"{name} fuse={fuse}: gates {ng} -> ops {tot}, passes {len(passes)}, switches {sum(p[1] for p in passes)}")
    names = [(0, "gate"), (28, "cgen"), (36, "diag1"), (40, "phase"), (41, "diaggen"), (42, "table"), (43, "table_reg"), (47, "pair"), (53, "twhad"), (57, "realph4"), (58, "twhad4"), (59, "switch"), (68, "?")]
    def nm(c):
        for (lo, n), (hi, _) in zip(names, names[1:]):
            if lo <= c < hi: return n
        return "?"
    for t, sw, nops, kinds in passes[:6]:
        agg = {}
        for c, k in kinds.items(): agg[nm(c)] = agg.get(nm(c), 0) + k
        print("   tile", t, "switches", sw, "ops", nops, agg)
    return passes

if __name__ == "__main__":
    for nm in sys.argv[1:] or ["qft30", "hea28", "random32"]:
        stats(nm, 1)
