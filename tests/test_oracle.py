"""The oracle against the reference's golden vector, hand-derived known answers and itself."""
import math

import numpy as np
import pytest

from damavand_b200 import circuits, gates as pgates
from oracle import oracle
from oracle.oracle import OracleCircuit


@pytest.mark.parametrize("method", ["multithreading", "brute_force"])
def test_reference_golden_vector(method):
    # the ONE known-answer test the reference holds: src/qubit_backend/circuit.rs:803-837
    c = OracleCircuit(2, method)
    c.add_hadamard_gate(0); c.add_hadamard_gate(1); c.add_cnot_gate(0, 1)
    c.forward()
    assert np.abs(c.amplitudes() - 0.5).max() < 1e-15


def test_golden_fixture_file():
    g = np.load("tests/golden/reference_h_h_cnot.npy")
    c = OracleCircuit(2); c.add_hadamard_gate(0); c.add_hadamard_gate(1); c.add_cnot_gate(0, 1); c.forward()
    assert np.abs(c.amplitudes() - g).max() < 1e-15


def test_known_answers():
    r = 1.0 / math.sqrt(2.0)
    c = OracleCircuit(1); c.add_hadamard_gate(0); c.forward()
    assert np.allclose(c.amplitudes(), [r, r], atol=1e-16)
    c = OracleCircuit(2); c.add_hadamard_gate(0); c.add_cnot_gate(0, 1); c.forward()
    assert np.allclose(c.amplitudes(), [r, 0, 0, r], atol=1e-16)          # Bell
    c = OracleCircuit(1); c.add_rotation_x_gate(0, math.pi); c.forward()
    assert np.allclose(c.amplitudes(), [0, -1j], atol=1e-15)
    th = 0.7
    c = OracleCircuit(1); c.add_hadamard_gate(0); c.add_rotation_z_gate(0, th); c.forward()
    assert np.allclose(c.amplitudes(), [r * np.exp(-0.5j * th), r * np.exp(0.5j * th)], atol=1e-15)
    for q in range(5):   # PauliX(q) moves the amplitude to index 1<<q: LSB = qubit 0
        c = OracleCircuit(5); c.add_pauli_x_gate(q, False); c.forward()
        assert c.amplitudes()[1 << q] == 1.0


def test_observables_are_skipped_and_forward_accumulates():
    c = OracleCircuit(3)
    c.add_hadamard_gate(0); c.add_pauli_z_gate(0, True); c.add_pauli_x_gate(1, True)
    c.forward()
    assert abs(c.amplitudes()[0] - 1 / math.sqrt(2)) < 1e-15 and c.amplitudes()[2] == 0
    c.forward()       # no implicit reset: H twice = identity
    assert abs(c.amplitudes()[0] - 1.0) < 1e-15


@pytest.mark.parametrize("n,seed", [(3, 1), (6, 2), (9, 3), (10, 4)])
def test_multithreading_equals_brute_force(n, seed):
    a, b = OracleCircuit(n, "multithreading"), OracleCircuit(n, "brute_force")
    circuits.random_circuit(a, n, 80, seed); circuits.random_circuit(b, n, 80, seed)
    a.forward(); b.forward()
    assert np.abs(a.amplitudes() - b.amplitudes()).max() < 1e-13
    assert abs(np.linalg.norm(a.amplitudes()) - 1) < 1e-13


def test_every_target_and_control_pair_vs_brute_force():
    n = 5
    rng = np.random.default_rng(0)
    for t in range(n):
        for c in [None] + [q for q in range(n) if q != t]:
            a, b = OracleCircuit(n, "multithreading"), OracleCircuit(n, "brute_force")
            for circ in (a, b):
                for q in range(n):
                    circ.add_rotation_y_gate(q, 0.3 + q); circ.add_rotation_z_gate(q, 0.2 * q + 0.1)
                if c is None:
                    circ.add_rotation_x_gate(t, 1.234)
                else:
                    circ.add_cnot_gate(c, t)
                circ.forward()
            assert np.abs(a.amplitudes() - b.amplitudes()).max() < 1e-14, (t, c)


def test_product_gate_matrices_match_oracle():
    th = 1.2345
    pairs = [("Hadamard", oracle.mat_hadamard(), None), ("PauliX", oracle.mat_pauli_x(), None),
             ("PauliY", oracle.mat_pauli_y(), None), ("PauliZ", oracle.mat_pauli_z(), None),
             ("RotationX", oracle.mat_rotation_x(th), th), ("RotationY", oracle.mat_rotation_y(th), th),
             ("RotationZ", oracle.mat_rotation_z(th), th), ("S", oracle.mat_s(), None), ("T", oracle.mat_t(), None),
             ("CNOT", oracle.mat_pauli_x(), None)]
    for name, om, p in pairs:
        pm = pgates.matrix(name, p)
        flat = [x for z in om for x in (z.real, z.imag)]
        assert pm == flat, name      # bit-identical


def test_set_parameters_and_reset():
    c = OracleCircuit(2)
    c.add_rotation_x_gate(0, 0.1); c.add_hadamard_gate(1); c.add_rotation_z_gate(1, 0.2); c.add_pauli_z_gate(0, True)
    c.set_parameters([1.0, 2.0, 3.0])
    assert c.gates[0].parameter == 1.0 and c.gates[2].parameter == 2.0
    c.reset()
    assert c.gates == [] and c.observables == [3]     # reset keeps observables (circuit.rs:303-306)


def test_samplers():
    c = OracleCircuit(12); circuits.layered(c, 12, 2); c.forward()
    p = c.measure_np()
    assert abs(p.sum() - 1) < 1e-12
    u = np.random.default_rng(1235).random(20000)
    s_seq = oracle.sample_sequential(p, u)
    s_faith = oracle.sample_sequential(p, u[:200], faithful=True)
    s_tree = oracle.sample_tree(p, u)
    assert (s_seq[:200] == s_faith).all()
    assert (s_seq == s_tree).all()
    # definition check against numpy: smallest k with inclusive prefix >= xsi
    cum = np.cumsum(p)
    k = np.searchsorted(cum, u * cum[-1], side="left")
    assert (k == s_seq).all()
    # empirical distribution
    hist = np.bincount(s_seq.astype(np.int64), minlength=p.size) / u.size
    assert np.abs(hist - p).max() < 0.02
    # edge cases: u = 0 -> index 0 ; deterministic state
    assert oracle.sample_sequential(p, np.array([0.0]))[0] == 0 and oracle.sample_tree(p, np.array([0.0]))[0] == 0
    d = np.zeros(16); d[5] = 1.0
    assert (oracle.sample_sequential(d, u[:50]) == 5).all() and (oracle.sample_tree(d, u[:50]) == 5).all()


def test_extract_expectation_values():
    c = OracleCircuit(4)
    c.add_pauli_z_gate(0, True); c.add_pauli_x_gate(2, True); c.add_pauli_y_gate(3, True); c.add_pauli_z_gate(1, False)
    out = c.extract_expectation_values([0b0000, 0b0001, 0b1100, 0b0110])
    assert out == [[1, 1, 1], [-1, 1, 1], [1, -1, -1], [1, -1, 1]]
    assert c.extract_expectation_values([]) == []


def test_distributed_sampler_semantics():
    c = OracleCircuit(10); circuits.layered(c, 10, 2); c.forward()
    p = c.measure_np()
    rng = np.random.default_rng(5)
    u1, u2 = rng.random(5000), rng.random(5000)
    for world in (2, 4, 8):
        s = oracle.sample_distributed(p, world, u1, u2, "sequential")
        t = oracle.sample_distributed(p, world, u1, u2, "tree")
        assert (s == t).all()
        hist = np.bincount(s.astype(np.int64), minlength=p.size) / u1.size
        assert np.abs(hist - p).max() < 0.03
    # world = 1 degenerates to the local sampler driven by the second draw
    assert (oracle.sample_distributed(p, 1, u1, u2) == oracle.sample_sequential(p, u2)).all()


def test_partner_rank_is_xor():
    L = oracle.lib()
    for per in (1, 4, 1024):
        for world_bits in (1, 2, 3):
            for j in range(world_bits):
                gap = per << j
                for r in range(1 << world_bits):
                    assert L.orc_compute_partner_rank(r, per, gap) == r ^ (1 << j)


# ---- committed fixtures (tests/golden/make_golden.py states where each one comes from) ----------------------------
def _golden():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join("tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("method", ["multithreading", "brute_force"])
def test_hand_derived_known_answers(method):
    mg = _golden()
    ka = np.load("tests/golden/known_answers.npz")
    assert set(ka.files) == set(mg.known_answers())
    for name in ka.files:
        c = mg.build_known(name, lambda n: OracleCircuit(n, method))
        c.forward()
        assert np.abs(c.amplitudes() - ka[name]).max() < 1e-15, name


@pytest.mark.parametrize("kind", ["layered", "hea", "qft", "random"])
def test_oracle_regression_fixtures(kind):
    """The oracle reproduces its own frozen outputs bit for bit (amplitudes, samples, expectation values)."""
    mg = _golden()
    fx = np.load("tests/golden/oracle_regression.npz")
    c = mg.regression_case(kind, OracleCircuit)
    c.forward()
    assert (c.amplitudes() == fx[f"{kind}_amplitudes"]).all()
    s = c.sample(64, uniforms=fx[f"{kind}_uniforms"], mode="sequential")
    assert (np.asarray(s, dtype=np.uint64) == fx[f"{kind}_samples"]).all()
    assert s == c.sample(64, uniforms=fx[f"{kind}_uniforms"], mode="tree")
    assert (np.asarray(c.extract_expectation_values(s)) == fx[f"{kind}_expectation"]).all()
    assert abs(np.linalg.norm(fx[f"{kind}_amplitudes"]) - 1.0) < 1e-13


def test_unitary_and_controlled_gates_match_brute_force():
    """The oracle's add_unitary_gate / add_controlled_gate (arbitrary 2x2 through the reference's update rule,
    circuit_multithreading.rs:9-54) against the brute_force restatement (full Kronecker operator)."""
    rng = np.random.default_rng(5)
    n = 6
    a, b = OracleCircuit(n, "multithreading"), OracleCircuit(n, "brute_force")
    for c in (a, b):
        r = np.random.default_rng(5)
        for q in range(n):
            c.add_hadamard_gate(q)
        for k in range(30):
            t = int(r.integers(0, n)); ctl = int(r.integers(0, n - 1)); ctl += ctl >= t
            z = r.normal(size=(2, 2)) + 1j * r.normal(size=(2, 2))
            u, _ = np.linalg.qr(z)
            if k % 2:
                c.add_controlled_gate(ctl, t, u.tolist())
            else:
                c.add_unitary_gate(t, u.tolist())
        c.forward()
    assert np.abs(a.amplitudes() - b.amplitudes()).max() < 1e-13
    assert abs(np.linalg.norm(a.amplitudes()) - 1.0) < 1e-12
