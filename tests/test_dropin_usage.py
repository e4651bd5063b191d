"""the call sequences the reference's users make, through the drop-in module name.

* the reference's own example (examples/single_node_single_gpu/single_node_single_gpu.py:7-20), verbatim
  apart from the injected uniforms;
* the sequence a PennyLane device plugin drives (docs/src/pennylane-plugin/README.md:23-37: a QNode of
  rotations returning <Z_w> for every wire).  pennylane and pennylane-damavand are not in the image, so the
  device below is a stand-in that maps PennyLane operation names onto the Circuit API the way a plugin does.
Both run on the GPU engine and on the oracle and must agree exactly (same seeded uniforms)."""
import numpy as np
import pytest

from oracle.oracle import OracleCircuit


def run_circuit(make, num_qubits, num_layers, uniforms):
    circuit = make(num_qubits)
    for l in range(num_layers):
        for i in range(num_qubits):
            circuit.add_rotation_x_gate(i, np.pi / (i + 1))
    for i in range(num_qubits):
        circuit.add_pauli_z_gate(i, True)
    circuit.forward()
    samples = circuit.sample(uniforms=uniforms)          # default 1000 shots
    return samples, np.mean(circuit.extract_expectation_values(samples), axis=0)


class QubitDeviceStandIn:
    """What a `damavand.qubit` PennyLane device does with a tape: reset, apply, sample, expval."""
    OPS = {"Hadamard": ("add_hadamard_gate", 0), "RX": ("add_rotation_x_gate", 1), "RY": ("add_rotation_y_gate", 1),
           "RZ": ("add_rotation_z_gate", 1), "CNOT": ("add_cnot_gate", 0)}
    PAULI = {"PauliX": "add_pauli_x_gate", "PauliY": "add_pauli_y_gate", "PauliZ": "add_pauli_z_gate"}

    def __init__(self, make, wires, shots=1000):
        self.circuit = make(wires)
        self.wires, self.shots = wires, shots

    def execute(self, operations, observables, uniforms):
        c = self.circuit
        c.reset()                                  # clears gates, keeps amplitudes reset (circuit.rs:303-306)
        c.observables = []                         # a plugin rebuilds its observables per tape
        for name, wires, params in operations:
            if name in self.PAULI:
                getattr(c, self.PAULI[name])(wires[0], False)
            else:
                meth, n_par = self.OPS[name]
                getattr(c, meth)(*wires, *params[:n_par])
        for name, wire in observables:
            getattr(c, self.PAULI[name])(wire, True)
        c.forward()
        samples = c.sample(self.shots, uniforms=uniforms)
        ev = np.asarray(c.extract_expectation_values(samples))
        return ev.mean(axis=0)


def tape(num_qubits, num_layers, theta):
    ops = []
    for l in range(num_layers):
        for w in range(num_qubits):
            ops.append(("RX", (w,), (theta[l, w, 0],)))
            ops.append(("RY", (w,), (theta[l, w, 1],)))
        for w in range(num_qubits - 1):
            ops.append(("CNOT", (w, w + 1), ()))
    ops.append(("Hadamard", (0,), ()))
    ops.append(("PauliX", (1,), ()))
    return ops, [("PauliZ", w) for w in range(num_qubits)]


def test_call_sequences_on_the_oracle_surface():
    """CPU: the sequences themselves are valid against the reference's API surface (oracle mirror)."""
    u = np.random.default_rng(5).random(1000)
    s, ev = run_circuit(OracleCircuit, 6, 5, u)
    assert len(s) == 1000 and ev.shape == (6,) and np.all(np.abs(ev) <= 1)
    dev = QubitDeviceStandIn(OracleCircuit, 5, shots=200)
    theta = np.random.default_rng(6).random((2, 5, 2))
    ops, obs = tape(5, 2, theta)
    r1 = dev.execute(ops, obs, np.random.default_rng(7).random(200))
    r2 = dev.execute(ops, obs, np.random.default_rng(7).random(200))     # a device is re-used across tapes
    assert r1.shape == (5,) and (r1 == r2).all()


@pytest.mark.gpu
@pytest.mark.parametrize("num_qubits", [2, 7, 14])
def test_reference_example_single_node_single_gpu(num_qubits):
    from damavand import Circuit          # the reference's module name
    u = np.random.default_rng(11).random(1000)
    sg, evg = run_circuit(lambda n: Circuit(n, apply_method="gpu"), num_qubits, 5, u)
    so, evo = run_circuit(OracleCircuit, num_qubits, 5, u)
    assert sg == so and (evg == evo).all()


@pytest.mark.gpu
def test_pennylane_style_device_gradient_loop():
    """Parameter-shift style loop: the same device object executes many tapes (reset + rebuild + forward)."""
    from damavand import Circuit
    n, layers = 13, 2
    g = QubitDeviceStandIn(lambda w: Circuit(w, "gpu"), n, shots=500)
    o = QubitDeviceStandIn(OracleCircuit, n, shots=500)
    rng = np.random.default_rng(12)
    theta = rng.random((layers, n, 2)) * 2 * np.pi
    for shift in (0.0, np.pi / 2, -np.pi / 2):
        th = theta.copy(); th[0, 3, 0] += shift
        ops, obs = tape(n, layers, th)
        u = np.random.default_rng(13).random(500)
        assert (g.execute(ops, obs, u) == o.execute(ops, obs, u)).all()
