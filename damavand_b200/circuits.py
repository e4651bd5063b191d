"""Seeded synthetic circuits of the shapes named in BASELINE.json `configs`.

Each generator only calls the reference's public ``Circuit.add_*`` methods
(/root/reference/src/qubit_backend/circuit.rs:608-691), so the same function drives the CUDA
``damavand_b200.Circuit`` and, in tests, the CPU oracle.  Angles are U[0, 2*pi) from
``numpy.random.default_rng(seed)`` drawn in gate order.  Returns the number of non-observable gates.
"""
from __future__ import annotations

import math

import numpy as np

TWO_PI = 2.0 * math.pi


def layered(circ, n: int, layers: int = 10, seed: int = 1234, observables: bool = True) -> int:
    """cfg 1: Hadamards, RX/RY/RZ on every qubit, CNOT ring; Z observable on every qubit."""
    rng = np.random.default_rng(seed)
    g = 0
    for _ in range(layers):
        for q in range(n):
            circ.add_hadamard_gate(q); g += 1
        for q in range(n):
            circ.add_rotation_x_gate(q, float(rng.random() * TWO_PI))
            circ.add_rotation_y_gate(q, float(rng.random() * TWO_PI))
            circ.add_rotation_z_gate(q, float(rng.random() * TWO_PI))
            g += 3
        if n > 1:
            for q in range(n):
                circ.add_cnot_gate(q, (q + 1) % n); g += 1
    if observables:
        for q in range(n):
            circ.add_pauli_z_gate(q, True)
    return g


def hea(circ, n: int, layers: int = 50, seed: int = 1234, observables: bool = False) -> int:
    """cfg 2 / cfg 5: hardware-efficient ansatz, RY+RZ on every qubit then a linear CNOT chain."""
    rng = np.random.default_rng(seed)
    g = 0
    for _ in range(layers):
        for q in range(n):
            circ.add_rotation_y_gate(q, float(rng.random() * TWO_PI))
            circ.add_rotation_z_gate(q, float(rng.random() * TWO_PI))
            g += 2
        for q in range(n - 1):
            circ.add_cnot_gate(q, q + 1); g += 1
    if observables:
        for q in range(n):
            circ.add_pauli_z_gate(q, True)
    return g


def qft_like(circ, n: int) -> int:
    """cfg 3: QFT-style circuit from API gates only.  Controlled-phase(phi) between c and t is
    RZ(phi/2)(t) CNOT(c,t) RZ(-phi/2)(t) CNOT(c,t) RZ(phi/2)(c) (up to a global phase);
    target-major order keeps consecutive gates on the same (high-stride) target."""
    g = 0
    for t in range(n - 1, -1, -1):
        circ.add_hadamard_gate(t); g += 1
        for c in range(t - 1, -1, -1):
            phi = math.pi / float(1 << (t - c))
            circ.add_rotation_z_gate(t, phi / 2.0)
            circ.add_cnot_gate(c, t)
            circ.add_rotation_z_gate(t, -phi / 2.0)
            circ.add_cnot_gate(c, t)
            circ.add_rotation_z_gate(c, phi / 2.0)
            g += 5
    return g


def random_circuit(circ, n: int, gates: int = 640, seed: int = 1234) -> int:
    """cfg 4: each gate uniform over {H, RX, RY, RZ, CNOT}; target uniform; control uniform != target."""
    rng = np.random.default_rng(seed)
    for _ in range(gates):
        kind = int(rng.integers(0, 5))
        t = int(rng.integers(0, n))
        if kind == 0:
            circ.add_hadamard_gate(t)
        elif kind == 1:
            circ.add_rotation_x_gate(t, float(rng.random() * TWO_PI))
        elif kind == 2:
            circ.add_rotation_y_gate(t, float(rng.random() * TWO_PI))
        elif kind == 3:
            circ.add_rotation_z_gate(t, float(rng.random() * TWO_PI))
        else:
            if n < 2:
                circ.add_hadamard_gate(t)
                continue
            c = int(rng.integers(0, n - 1))
            if c >= t:
                c += 1
            circ.add_cnot_gate(c, t)
    return gates


def workload(name: str):
    """Resolve a workload name: the five BASELINE configs, or `qft<n>`, `hea<n>[x<layers>]`,
    `random<n>[x<gates>]`, `layered<n>[x<layers>]` for other sizes.  Returns (n_qubits, builder)."""
    import re
    m = re.fullmatch(r"(qft|hea|random|layered|hhi)(\d+)(?:x(\d+))?", name)
    if not m:
        raise KeyError(name)
    kind, n, k = m.group(1), int(m.group(2)), m.group(3)
    if kind == "qft":
        return n, (lambda c: qft_like(c, n))
    if kind == "hhi":      # one Hadamard on each of the k highest qubits: ONE pass over a high-stride tile, next to no arithmetic
        top = int(k) if k else 9
        def build(c):
            for q in range(n - top, n):
                c.add_hadamard_gate(q)
            return top
        return n, build
    if kind == "hea":
        layers = int(k) if k else (50 if n <= 28 else 10)
        return n, (lambda c: hea(c, n, layers, observables=(n >= 34)))
    if kind == "random":
        g = int(k) if k else 640
        return n, (lambda c: random_circuit(c, n, g))
    layers = int(k) if k else 10
    return n, (lambda c: layered(c, n, layers))


WORKLOADS = {
    "layered20": lambda c: layered(c, 20, 10),
    "hea28": lambda c: hea(c, 28, 50),
    "qft30": lambda c: qft_like(c, 30),
    "random32": lambda c: random_circuit(c, 32, 640),
    "hea34": lambda c: hea(c, 34, 10, observables=True),
}
