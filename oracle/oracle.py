"""CPU oracle for the damavand statevector hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
``damavand_b200`` never does.

It restates, with file:line citations into the reference (``/root/reference``):

* gate matrices                      -- src/qubit_backend/gates.rs (constructors)
* ``multithreading`` apply method     -- src/qubit_backend/circuit_multithreading.rs:9-54  (oracle.c)
* ``brute_force`` apply method        -- src/qubit_backend/circuit_brute_force.rs:9-12, src/utils.rs:32-45,119-245 (numpy)
* forward / measure / sample / extract_expectation_values / set_parameters / reset
                                      -- src/qubit_backend/circuit.rs
* distributed sampling               -- src/qubit_backend/circuit_distributed.rs:42-129

Parity status: the reference cannot be built in this image (no cargo / MPI).  The restatement is
pinned on the one golden vector the reference's tests hold (circuit.rs:803-837) and the two
restated methods are cross-checked against each other; otherwise PARITY UNPINNED.

The reference draws uniforms from ``thread_rng()``; here they are injected so that both sides of a
parity test consume the same draws.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc, no FMA contraction, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _LIB_PATH, src]
        )
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.orc_apply_gate.argtypes = [dp, dp, ctypes.c_uint64, dp, ctypes.c_int, ctypes.c_int]
        L.orc_apply_gate.restype = None
        L.orc_measure.argtypes = [dp, ctypes.c_uint64, dp]
        L.orc_measure.restype = None
        L.orc_sample_sequential.argtypes = [dp, ctypes.c_uint64, dp, ctypes.c_uint64, u64p, ctypes.c_int]
        L.orc_sample_sequential.restype = None
        L.orc_sample_tree.argtypes = [dp, ctypes.c_uint64, dp, ctypes.c_uint64, u64p]
        L.orc_sample_tree.restype = None
        L.orc_tree_total.argtypes = [dp, ctypes.c_uint64]
        L.orc_tree_total.restype = ctypes.c_double
        L.orc_extract_expectation_values.argtypes = [u64p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int), ctypes.c_int, dp]
        L.orc_extract_expectation_values.restype = None
        L.orc_compute_partner_rank.argtypes = [ctypes.c_uint64] * 3
        L.orc_compute_partner_rank.restype = ctypes.c_uint64
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------------------------
# Gate matrices, exactly as src/qubit_backend/gates.rs writes them (libm cos/sin/sqrt via `math`).
# Each returns the row-major 2x2 as a list of 4 Python complex numbers [m00, m01, m10, m11].
# --------------------------------------------------------------------------------------------
def mat_hadamard():  # gates.rs:97-125  (1./(2.0).sqrt(), not FRAC_1_SQRT_2)
    h = 1.0 / math.sqrt(2.0)
    return [complex(h, 0.0), complex(h, 0.0), complex(h, 0.0), complex(-1.0 / math.sqrt(2.0), 0.0)]


def mat_pauli_x():  # gates.rs:223-240 ; CNOT uses the same 2x2, gates.rs:169-181
    return [0j, 1 + 0j, 1 + 0j, 0j]


def mat_pauli_y():  # gates.rs:288-304
    return [0j, complex(0.0, -1.0), complex(0.0, 1.0), 0j]


def mat_pauli_z():  # gates.rs:352-368
    return [1 + 0j, 0j, 0j, complex(-1.0, 0.0)]


def mat_rotation_x(theta):  # gates.rs:417-446
    c, s = math.cos(theta / 2.0), math.sin(theta / 2.0)
    return [complex(c, 0.0), complex(0.0, -s), complex(0.0, -s), complex(c, 0.0)]


def mat_rotation_y(theta):  # gates.rs:489-518
    c, s = math.cos(theta / 2.0), math.sin(theta / 2.0)
    return [complex(c, 0.0), complex(-s, 0.0), complex(s, 0.0), complex(c, 0.0)]


def mat_rotation_z(theta):  # gates.rs:561-584
    return [
        complex(math.cos(-theta / 2.0), math.sin(-theta / 2.0)), 0j,
        0j, complex(math.cos(theta / 2.0), math.sin(theta / 2.0)),
    ]


def mat_s():  # gates.rs:626-641
    return [1 + 0j, 0j, 0j, complex(0.0, 1.0)]


def mat_t():  # gates.rs:684-703
    return [1 + 0j, 0j, 0j, complex(math.sqrt(2.0) / 2.0, math.sqrt(2.0) / 2.0)]


class _Gate:
    __slots__ = ("name", "target", "control", "parameter", "custom")

    def __init__(self, name, target, control=None, parameter=None, custom=None):
        self.name, self.target, self.control, self.parameter, self.custom = name, target, control, parameter, custom

    def matrix(self):
        n = self.name
        if n == "Unitary":      # caller-supplied 2x2, row-major (the product's add_unitary_gate / add_controlled_gate)
            return list(self.custom)
        if n == "Hadamard":
            return mat_hadamard()
        if n in ("PauliX", "CNOT"):
            return mat_pauli_x()
        if n == "PauliY":
            return mat_pauli_y()
        if n == "PauliZ":
            return mat_pauli_z()
        if n == "RotationX":
            return mat_rotation_x(self.parameter)
        if n == "RotationY":
            return mat_rotation_y(self.parameter)
        if n == "RotationZ":
            return mat_rotation_z(self.parameter)
        if n == "S":
            return mat_s()
        if n == "T":
            return mat_t()
        raise ValueError(n)


def _flat(m):
    out = np.empty(8, dtype=np.float64)
    for k, z in enumerate(m):
        out[2 * k], out[2 * k + 1] = z.real, z.imag
    return out


# --------------------------------------------------------------------------------------------
# brute_force restatement: full 2^n x 2^n operator (utils.rs:119-245), qubit 0 = last Kronecker
# factor (utils.rs:199-205 iterates qubits in reverse).
# --------------------------------------------------------------------------------------------
def brute_force_operator(num_qubits: int, gate: _Gate) -> np.ndarray:
    I2 = np.eye(2, dtype=np.complex128)
    m = np.array(gate.matrix(), dtype=np.complex128).reshape(2, 2)
    if gate.control is None:
        if num_qubits == 1:
            return m
        mats = [m if q == gate.target else I2 for q in reversed(range(num_qubits))]
        op = mats[0]
        for x in mats[1:]:
            op = np.kron(op, x)
        return op
    p0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    p1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    sx = m     # CNOT: the Pauli-X block (utils.rs:230-241); a general controlled gate puts its own 2x2 there (extension)
    inactive, active = [], []
    for q in reversed(range(num_qubits)):  # utils.rs:230-241
        if q == gate.control:
            inactive.append(p0), active.append(p1)
        elif q == gate.target:
            inactive.append(I2), active.append(sx)
        else:
            inactive.append(I2), active.append(I2)
    a, b = inactive[0], active[0]
    for x in inactive[1:]:
        a = np.kron(a, x)
    for x in active[1:]:
        b = np.kron(b, x)
    return a + b


# --------------------------------------------------------------------------------------------
# Samplers on a probability vector
# --------------------------------------------------------------------------------------------
def sample_sequential(probs: np.ndarray, u: np.ndarray, faithful: bool = False) -> np.ndarray:
    """utils.rs:258-277 with injected uniforms (reference semantics, sequential cumulative)."""
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.shape[0], dtype=np.uint64)
    lib().orc_sample_sequential(_dp(probs), probs.shape[0], _dp(u), u.shape[0],
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), int(faithful))
    return out


def sample_tree(probs: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Pairwise-tree summation order = the CUDA sampler's specified order (see oracle.c)."""
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = probs.shape[0]
    assert n & (n - 1) == 0
    out = np.empty(u.shape[0], dtype=np.uint64)
    lib().orc_sample_tree(_dp(probs), n, _dp(u), u.shape[0],
                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    return out


def tree_total(probs: np.ndarray) -> float:
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    return float(lib().orc_tree_total(_dp(probs), probs.shape[0]))


def sample_distributed(probs: np.ndarray, world: int, u_rank: np.ndarray, u_local: np.ndarray,
                       mode: str = "sequential") -> np.ndarray:
    """circuit_distributed.rs:42-129 with injected uniforms.

    Rank r owns probs[r*per:(r+1)*per].  Per shot s: the root picks a rank from the per-rank totals
    with u_rank[s] (:83-91, sample_from_discrete_distribution over the totals), then that rank draws
    a local index with a second, independent draw u_local[s] (:99-112) and reports
    local + per*rank (:113-115).  mode selects the summation order of the *local* sums
    ("sequential" = reference, "tree" = the CUDA sampler's documented order).
    """
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    per = probs.shape[0] // world
    chunks = [probs[r * per:(r + 1) * per] for r in range(world)]
    if mode == "sequential":
        # cumulative_distribution.last(), :54-58; np.cumsum (add.accumulate) is a strict
        # left-to-right fp64 sum, same as that loop (0.0 + p0 == p0 exactly).
        totals = [float(np.cumsum(c)[-1]) for c in chunks]
    else:
        totals = [tree_total(c) for c in chunks]
    totals = np.array(totals, dtype=np.float64)
    ranks = sample_sequential(totals, u_rank)  # the rank draw is always the sequential rule
    out = np.empty(u_rank.shape[0], dtype=np.uint64)
    for r in range(world):
        sel = np.nonzero(ranks == r)[0]
        if sel.size == 0:
            continue
        loc = sample_sequential(chunks[r], u_local[sel]) if mode == "sequential" else sample_tree(chunks[r], u_local[sel])
        out[sel] = loc + np.uint64(per * r)
    return out


# --------------------------------------------------------------------------------------------
# OracleCircuit: the reference's Circuit API (circuit.rs:78-770) on the CPU restatement.
# --------------------------------------------------------------------------------------------
class OracleCircuit:
    METHODS = ("multithreading", "brute_force")

    def __init__(self, num_qubits: int, apply_method: Optional[str] = None):
        method = "multithreading" if apply_method is None else apply_method  # circuit.rs:95
        if method not in self.METHODS:
            raise ValueError(f"oracle implements only {self.METHODS}, got {method}")
        self.num_qubits = num_qubits
        self.apply_method = method
        self.gates: List[_Gate] = []
        self.observables: List[int] = []
        self.reset_amplitudes()

    # -- state lifecycle ---------------------------------------------------------------------
    def reset_amplitudes(self):  # circuit.rs:262-302 / :164-175
        n = 1 << self.num_qubits
        self.state = np.zeros(2 * n, dtype=np.float64)
        self.state[0] = 1.0
        self._scratch = None

    def reset(self):  # circuit.rs:303-306 (observables are NOT cleared)
        self.reset_amplitudes()
        self.gates = []

    def set_parameters(self, parameters: Sequence[float]):  # circuit.rs:308-322
        idx = [i for i, g in enumerate(self.gates) if g.parameter is not None]
        for i, p in zip(idx, parameters):
            self.gates[i].parameter = float(p)

    # -- gate list ---------------------------------------------------------------------------
    def add_hadamard_gate(self, q):
        self.gates.append(_Gate("Hadamard", q))

    def add_rotation_x_gate(self, q, theta):
        self.gates.append(_Gate("RotationX", q, parameter=float(theta)))

    def add_rotation_y_gate(self, q, theta):
        self.gates.append(_Gate("RotationY", q, parameter=float(theta)))

    def add_rotation_z_gate(self, q, theta):
        self.gates.append(_Gate("RotationZ", q, parameter=float(theta)))

    def _pauli(self, name, q, is_observable):
        self.gates.append(_Gate(name, q))
        if is_observable:
            self.observables.append(len(self.gates) - 1)  # circuit.rs:652-654

    def add_pauli_x_gate(self, q, is_observable):
        self._pauli("PauliX", q, is_observable)

    def add_pauli_y_gate(self, q, is_observable):
        self._pauli("PauliY", q, is_observable)

    def add_pauli_z_gate(self, q, is_observable):
        self._pauli("PauliZ", q, is_observable)

    def add_cnot_gate(self, control, target):
        self.gates.append(_Gate("CNOT", target, control=control))

    # gates.rs defines S (:617-641) and T (:675-703) but circuit.rs has no add_ method for them; the product offers
    # add_s_gate / add_t_gate as extensions and the oracle mirrors that with the reference's matrices
    def add_unitary_gate(self, q, matrix):
        self.gates.append(_Gate("Unitary", q, custom=[complex(z) for row in matrix for z in row]))

    def add_controlled_gate(self, control, target, matrix):   # update rule unchanged: circuit_multithreading.rs:36-38
        self.gates.append(_Gate("Unitary", target, control=control, custom=[complex(z) for row in matrix for z in row]))

    def add_s_gate(self, q):
        self.gates.append(_Gate("S", q))

    def add_t_gate(self, q):
        self.gates.append(_Gate("T", q))

    # -- forward -----------------------------------------------------------------------------
    def forward(self, max_gates: Optional[int] = None):  # circuit.rs:341-375 (no implicit reset)
        n = 1 << self.num_qubits
        applied = 0
        for gi, g in enumerate(self.gates):
            if gi in self.observables:  # :347-349
                continue
            if max_gates is not None and applied >= max_gates:
                break
            if self.apply_method == "multithreading":
                if self._scratch is None:
                    self._scratch = np.empty(2 * n, dtype=np.float64)
                m = _flat(g.matrix())
                lib().orc_apply_gate(_dp(self.state), _dp(self._scratch), n, _dp(m),
                                     -1 if g.control is None else int(g.control), int(g.target))
            else:
                op = brute_force_operator(self.num_qubits, g)
                z = op.dot(self.state.view(np.complex128))
                self.state = np.ascontiguousarray(z).view(np.float64).copy()
            applied += 1
        return applied

    def retrieve_amplitudes_on_host(self):
        pass

    # -- observation -------------------------------------------------------------------------
    def amplitudes(self) -> np.ndarray:
        return self.state.view(np.complex128)

    def get_real_part_state(self):  # circuit.rs:596-598
        return self.state[0::2].tolist()

    def get_imaginary_part_state(self):  # circuit.rs:600-602
        return self.state[1::2].tolist()

    def measure_np(self) -> np.ndarray:  # circuit.rs:565-589
        n = 1 << self.num_qubits
        p = np.empty(n, dtype=np.float64)
        lib().orc_measure(_dp(self.state), n, _dp(p))
        return p

    def measure(self):
        return self.measure_np().tolist()

    def sample(self, num_samples: Optional[int] = None, uniforms=None, mode: str = "sequential",
               faithful: bool = False):  # circuit.rs:434-458, default 1000 shots
        shots = 1000 if num_samples is None else num_samples
        if uniforms is None:
            uniforms = np.random.default_rng().random(shots)
        u = np.asarray(uniforms, dtype=np.float64)[:shots]
        p = self.measure_np()
        s = sample_sequential(p, u, faithful) if mode == "sequential" else sample_tree(p, u)
        return [int(x) for x in s]

    def extract_expectation_values(self, samples):  # circuit.rs:494-513
        s = np.ascontiguousarray(np.asarray(samples, dtype=np.uint64))
        q = np.ascontiguousarray(np.array([self.gates[i].target for i in self.observables], dtype=np.int32))
        out = np.empty((s.shape[0], q.shape[0]), dtype=np.float64)
        if q.shape[0] and s.shape[0]:
            lib().orc_extract_expectation_values(
                s.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), s.shape[0],
                q.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), q.shape[0], _dp(out))
        return out.tolist()

    def get_fidelity_between_two_states_with_parameters(self, p1, p2):  # circuit.rs:753-769
        self.reset_amplitudes(); self.set_parameters(p1); self.forward()
        s1 = self.amplitudes().copy()
        self.reset_amplitudes(); self.set_parameters(p2); self.forward()
        s2 = self.amplitudes().copy()
        f = abs(np.dot(np.conj(s1), s2)) ** 2  # circuit_metrics.rs:24
        self.reset()                             # circuit_metrics.rs:30
        return float(f)
