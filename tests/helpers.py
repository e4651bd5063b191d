"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import ctypes
import os
import subprocess

import numpy as np

from damavand_b200 import _lib, gates as pgates
from oracle.oracle import OracleCircuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libdvd_emu.so")
_emu = None


class Recorder:
    """Records add_* calls so the same circuit can be replayed on several backends."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name.startswith("add_"):
            def f(*a):
                self.calls.append((name, a))
            return f
        raise AttributeError(name)

    def replay(self, circ):
        for name, a in self.calls:
            getattr(circ, name)(*a)
        return circ


def gate_array(circ):
    """(ctypes Gate array, n) of a circuit object exposing .gates/.observables in oracle format."""
    obs = set(circ.observables)
    todo = [g for i, g in enumerate(circ.gates) if i not in obs]
    arr = (_lib.Gate * max(1, len(todo)))()
    for k, g in enumerate(todo):
        arr[k].target = g.target
        arr[k].control = -1 if g.control is None else g.control
        arr[k].m[:] = pgates.matrix(g.name, g.parameter, None if g.custom is None else pgates.from_2x2([g.custom[:2], g.custom[2:]]))
    return arr, len(todo)


def emu():
    global _emu
    if _emu is None:
        src = os.path.join(EMU_DIR, "emu.cpp")
        deps = [src] + [os.path.join(ROOT, "damavand_b200", "csrc", f) for f in ("planner.cpp", "planner.h", "tile_core.cuh")]
        if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", EMU_LIB, src])
        L = ctypes.CDLL(EMU_LIB)
        L.emu_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.Gate), ctypes.c_int64, ctypes.c_int,
                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
        L.emu_run.restype = ctypes.c_int
        L.emu_run_sparse.argtypes = L.emu_run.argtypes
        L.emu_run_sparse.restype = ctypes.c_int
        L.emu_error.restype = ctypes.c_char_p
        L.emu_max_bank_conflict.restype = ctypes.c_int
        L.emu_set_fused.argtypes = [ctypes.c_int]
        L.emu_set_store.argtypes = [ctypes.c_int]
        L.emu_set_defer.argtypes = [ctypes.c_int]
        L.emu_last_defer.restype = ctypes.c_int
        L.emu_last_store.restype = ctypes.c_int
        L.emu_plan_replay_check.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.Gate), ctypes.POINTER(_lib.Gate), ctypes.c_int64,
                                            ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.emu_plan_replay_check.restype = ctypes.c_int
        _emu = L
    return _emu


def emu_run(oracle_circ: OracleCircuit, world: int = 1, state=None, fuse: bool = True, track_support: bool = False,
            fused_remap: bool = True, store_side: int = 2, defer: int = -1):
    """Run the recorded gates of `oracle_circ` through the CPU replay of the CUDA path.
    fused_remap: global<->local swaps ride on the next pass's load (the engine's default when both chunks fit), else
    every swap is an exchange of its own (in-place peer swap / staged NCCL path).
    store_side: 0 = swaps ride on loads only, 1 = the swaps that end the schedule (layout restore) ride on the STORE of the
    last gate pass when they can, 2 = every swap round rides on a store where it can (the engine's default).
    defer: tail-deferral threshold of the distributed schedule (-1: the planner picks the cheapest of a few).
    track_support: replay the engine's support tracking after a reset -- only amplitude 0 of every rank's chunk is
    stored, the rest of the buffer is NaN (never-written memory) and must never be read."""
    n = oracle_circ.num_qubits
    arr, ng = gate_array(oracle_circ)
    if track_support:
        assert state is None
        n_local = n - (world.bit_length() - 1)
        # chunks below one tile are reset with a real memset (engine.cu set_zero_state)
        state = np.full(2 << n, np.nan if n_local >= 12 else 0.0, dtype=np.float64)
        chunk = (2 << n) // world
        for r in range(world):
            state[r * chunk:r * chunk + 2] = 0.0
        state[0] = 1.0
    elif state is None:
        state = np.zeros(2 << n, dtype=np.float64)
        state[0] = 1.0
    stats = (ctypes.c_int64 * 4)()
    run = emu().emu_run_sparse if track_support else emu().emu_run
    emu().emu_set_fused(int(fused_remap))
    emu().emu_set_store(int(store_side))
    emu().emu_set_defer(int(defer))
    rc = run(n, world, arr, ng, int(fuse), state.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stats)
    if rc != 0:
        raise RuntimeError(emu().emu_error().decode())
    return state.view(np.complex128), dict(passes=stats[0], swaps=stats[1], switches=stats[2], ops=stats[3],
                                           defer=emu().emu_last_defer(), store_side=emu().emu_last_store())


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| relative to the LARGEST magnitude of the reference state: the reading of "1e-12 relative" for the
    amplitudes of a normalised state (an amplitude that cancels to ~0 has no meaningful relative error of its own).
    `l2_err` is the second, norm-wise metric; the parity tests at config scale assert both."""
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)


def l2_err(a: np.ndarray, b: np.ndarray) -> float:
    """||a - b||_2 / ||b||_2."""
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
