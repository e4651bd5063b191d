"""The last GPU seconds of round 2 (1 GPU): smoke() and a variational loop -- the same circuit flushed with new angles, so
the plan cache misses and the recorded planner choices are replayed (planner.h: ChoiceMemoTable) -- against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.smoke()
from damavand_b200 import Circuit, circuits
from oracle.oracle import OracleCircuit
n = 18
c, o = Circuit(n, "gpu"), OracleCircuit(n)
for x in (c, o):
    circuits.hea(x, n, 4)
n_par = sum(1 for gt in o.gates if gt.parameter is not None)
for it in range(4):
    params = [0.1 + 0.01 * it + 0.003 * k for k in range(n_par)]
    t0 = time.perf_counter()
    for x in (c, o):
        x.set_parameters(params); x.reset_amplitudes()
    c.forward(); a = c.state_numpy()
    dt = time.perf_counter() - t0
    o.forward(); b = o.amplitudes()
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 1e-12, err
    print(f"iteration {it}: rel_err {err:.1e}, plan_cache_hits {c.stats()['plan_cache_hits']}, reset+forward+readback {dt*1e3:.1f} ms")
print("variational loop ok")
