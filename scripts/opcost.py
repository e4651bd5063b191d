"""Marginal cost of each op class of k_tile_pass on the GPU (development microbenchmark).

Builds one-pass circuits through the public API whose op lists are k copies of one op class and
prints ms per pass and SM cycles per op per CTA pair.  Usage: python scripts/opcost.py [n_qubits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from damavand_b200 import Circuit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
waves = (1 << (n - 12)) / 296.0


def time_circuit(build, reps=5):
    c = Circuit(n, "gpu")
    c.add_hadamard_gate(n - 1); c.forward(); c.gates.clear()
    build(c)
    c.forward_async(); c.synchronize()         # warm-up (plan + launch)
    c.stats_reset()
    c.timer_begin()
    for _ in range(reps):
        c.forward_async()
    ms = c.timer_end() / reps
    st = c.stats()
    c.close()
    return ms, st["tile_passes"] // reps, st["stage_switches"] // reps


def report(name, ks, build):
    res = []
    for k in ks:
        ms, passes, sw = time_circuit(lambda c: build(c, k))
        res.append((k, ms, passes, sw))
    base = res[0]
    line = f"{name:28s}"
    for k, ms, passes, sw in res:
        line += f" k={k}: {ms:7.3f} ms ({passes}p,{sw}s)"
    k0, m0 = res[0][0], res[0][1]
    k1, m1 = res[-1][0], res[-1][1]
    cyc = (m1 - m0) / (k1 - k0) * 1e-3 * 1.965e9 / waves
    print(line + f"  -> {cyc:7.0f} cycles/op/wave", flush=True)


rng = np.random.default_rng(0)
q = 5          # first target: lands in the IO register group, no stage switch needed
ks = [1, 16, 64]
report("RY(q) real gate", ks, lambda c, k: [c.add_rotation_y_gate(q, 0.1 + 0.01 * i) for i in range(k)])
report("H(q) hadamard", ks, lambda c, k: [c.add_hadamard_gate(q) for i in range(k)])
report("RX(q) rx-like gate", ks, lambda c, k: [c.add_rotation_x_gate(q, 0.1 + 0.01 * i) for i in range(k)])
report("RY(q) RZ(q) gate+diag1", ks, lambda c, k: [(c.add_rotation_y_gate(q, 0.1 + 0.01 * i), c.add_rotation_z_gate(q, 0.3)) for i in range(k)])
report("RY(q) RX(q) real+rx", ks, lambda c, k: [(c.add_rotation_y_gate(q, 0.1 + 0.01 * i), c.add_rotation_x_gate(q, 0.3)) for i in range(k)])
# two targets in different register groups, made non-commuting by a CNOT between them: forces switches
a, b = 5, 9
report("RY(a) CX(a,b) RY(b) CX(b,a)", [1, 8, 32], lambda c, k: [(c.add_rotation_y_gate(a, 0.1), c.add_cnot_gate(a, b), c.add_rotation_y_gate(b, 0.2), c.add_cnot_gate(b, a)) for i in range(k)])
# thread-level controlled phase pairs -> table ops
report("CZ-like RZ CX RZ CX (t=5,c=20)", [1, 8, 32], lambda c, k: [(c.add_rotation_y_gate(5, 0.1), c.add_rotation_z_gate(5, 0.2), c.add_cnot_gate(20, 5), c.add_rotation_z_gate(5, -0.2), c.add_cnot_gate(20, 5)) for i in range(k)])
