"""Gate matrices with the reference's exact formulas (/root/reference/src/qubit_backend/gates.rs).

Each function returns the row-major 2x2 as 8 doubles
``[m00.re, m00.im, m01.re, m01.im, m10.re, m10.im, m11.re, m11.im]`` -- the layout the C ABI takes.
libm ``cos``/``sin``/``sqrt`` through ``math`` match what Rust's f64 methods call.
"""
from __future__ import annotations

import math


def hadamard():  # gates.rs:97-125: 1./(2.0 as f64).sqrt(), and -1./(2.0).sqrt() for m11
    h = 1.0 / math.sqrt(2.0)
    return [h, 0.0, h, 0.0, h, 0.0, -1.0 / math.sqrt(2.0), 0.0]


def pauli_x():  # gates.rs:223-240 (CNOT carries the same 2x2, gates.rs:169-181)
    return [0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0]


def pauli_y():  # gates.rs:288-304
    return [0.0, 0.0, 0.0, -1.0, 0.0, 1.0, 0.0, 0.0]


def pauli_z():  # gates.rs:352-368
    return [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0]


def rotation_x(theta: float):  # gates.rs:417-446
    c, s = math.cos(theta / 2.0), math.sin(theta / 2.0)
    return [c, 0.0, 0.0, -s, 0.0, -s, c, 0.0]


def rotation_y(theta: float):  # gates.rs:489-518
    c, s = math.cos(theta / 2.0), math.sin(theta / 2.0)
    return [c, 0.0, -s, 0.0, s, 0.0, c, 0.0]


def rotation_z(theta: float):  # gates.rs:561-584
    return [math.cos(-theta / 2.0), math.sin(-theta / 2.0), 0.0, 0.0,
            0.0, 0.0, math.cos(theta / 2.0), math.sin(theta / 2.0)]


def s_gate():  # gates.rs:626-641
    return [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]


def t_gate():  # gates.rs:684-703
    r = math.sqrt(2.0) / 2.0
    return [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, r, r]


_BUILDERS = {
    "Hadamard": lambda p: hadamard(),
    "CNOT": lambda p: pauli_x(),
    "PauliX": lambda p: pauli_x(),
    "PauliY": lambda p: pauli_y(),
    "PauliZ": lambda p: pauli_z(),
    "RotationX": rotation_x,
    "RotationY": rotation_y,
    "RotationZ": rotation_z,
    "S": lambda p: s_gate(),
    "T": lambda p: t_gate(),
}


def from_2x2(m):
    """Row-major 2x2 of complex numbers (nested or flat, anything numpy-like) -> the 8 doubles of the C ABI."""
    flat = [complex(z) for row in m for z in (row if hasattr(row, "__len__") else [row])]
    if len(flat) != 4:
        raise ValueError("expected a 2x2 matrix")
    out = []
    for z in flat:
        out += [float(z.real), float(z.imag)]
    return out


def matrix(name: str, parameter=None, custom=None):
    if name == "Unitary":      # caller-supplied 2x2 (add_unitary_gate / add_controlled_gate)
        return list(custom)
    return _BUILDERS[name](parameter)
