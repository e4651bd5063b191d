"""``Circuit`` -- the reference's Python-facing class for the `gpu` / `distributed_gpu` apply methods.

Mirrors the PyO3 class of /root/reference/src/qubit_backend/circuit.rs:60-770 method for method
(same names, argument order, defaults and return shapes), so user code and the pennylane-damavand
device keep working.  Everything numerical happens in libdamavand_b200.so through the C ABI
(include/damavand_b200.h); there is no CPU path here: the CPU apply methods of the reference
(`brute_force`, `shuffle`, `multithreading`, `distributed_cpu`) are not part of this package and
raise ``NotImplementedError``.

Deliberate, documented differences:
  * ``forward()`` does not copy the state to the host (the reference does, circuit.rs:373-374);
    ``get_real_part_state`` / ``get_imaginary_part_state`` / ``retrieve_amplitudes_on_host`` fetch it
    on demand.
  * ``sample(..., uniforms=...)`` optionally takes the uniform draws (the reference uses an
    unseedable ``thread_rng``, utils.rs:259); prefix sums use the pairwise-tree order.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib, distributed, gates

_GPU_METHODS = ("gpu", "distributed_gpu")
_CPU_METHODS = ("brute_force", "shuffle", "multithreading", "distributed_cpu")


def _dp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class _Profiler:
    """src/profiler.rs:11-43: iterations + cumulated elapsed nanoseconds, mean = cumulated / iterations.  The reference
    brackets host code with Instant::now(); here the buckets are fed with DEVICE time where the work is on the device:
    Forward = CUDA events on the engine's stream around the fused passes of one forward(), InterGPUCommunication = CUDA
    events around the passes whose load carries a global<->local qubit swap over NVLink (plus any stand-alone exchange)
    of that forward (csrc/engine.cu: remap_ms / swap_ms), Sampling = wall clock of sample() (host buffers in and out).
    InterNodeCommunication stays empty: one node (the reference feeds it from its MPI exchange, circuit_distributed_gpu.rs:109-115)."""

    def __init__(self):
        self.iterations = 0
        self.cumulated = 0.0
        self._t0 = 0.0

    def start(self):
        self._t0 = time.perf_counter_ns()

    def stop(self):
        self.cumulated += time.perf_counter_ns() - self._t0
        self.iterations += 1

    def add_ms(self, ms: float):
        self.cumulated += ms * 1e6
        self.iterations += 1

    def mean(self):
        return self.cumulated / self.iterations if self.iterations else 0.0


class Circuit:
    def __init__(self, num_qubits: int, apply_method: Optional[str] = None):
        method = "multithreading" if apply_method is None else apply_method   # circuit.rs:95
        if method in _CPU_METHODS:
            raise NotImplementedError(
                f"apply_method={method!r} is one of the reference's CPU methods; damavand_b200 implements only "
                f"{_GPU_METHODS} (there is deliberately no CPU fallback)")
        if method not in _GPU_METHODS:
            raise ValueError(f"Apply method not recognized: {method}")       # circuit.rs:112
        if num_qubits < 1:
            raise ValueError("num_qubits must be >= 1")
        self._lib = _lib.load()
        self.num_qubits = int(num_qubits)
        self.apply_method = method
        self.gates: List[list] = []        # [name, target, control, parameter]
        self.observables: List[int] = []
        self._gate_cache = None            # (contents key, marshalled dvd_gate array) of the last forward
        self._handle = ctypes.c_void_p()
        self._profilers = {k: _Profiler() for k in ("forward", "inter_node", "inter_gpu", "sampling")}
        if self._lib.dvd_device_count() == 0:
            raise _lib.DamavandError("Could not find any GPU.")              # circuit.rs:143-145
        if method == "gpu":
            self.rank, self.num_nodes = 0, 1
            device = int(os.environ.get("DVD_DEVICE", distributed.local_device() if "LOCAL_RANK" in os.environ else 0))
            _lib.check(self._lib.dvd_create(self.num_qubits, device, ctypes.byref(self._handle)), "dvd_create")
        else:
            self.rank, self.num_nodes = distributed.initialize()
            device = distributed.local_device() % max(1, self._lib.dvd_device_count())
            if self.num_nodes == 1:
                _lib.check(self._lib.dvd_create(self.num_qubits, device, ctypes.byref(self._handle)), "dvd_create")
            else:
                buf = (ctypes.c_ubyte * _lib.NCCL_ID_BYTES)()
                if self.rank == 0:
                    _lib.check(self._lib.dvd_nccl_unique_id(buf), "dvd_nccl_unique_id")
                ident = distributed.broadcast_bytes(bytes(buf) if self.rank == 0 else None, _lib.NCCL_ID_BYTES)
                idbuf = (ctypes.c_ubyte * _lib.NCCL_ID_BYTES).from_buffer_copy(ident)
                _lib.check(self._lib.dvd_create_distributed(self.num_qubits, device, self.rank, self.num_nodes,
                                                            idbuf, ctypes.byref(self._handle)),
                           "dvd_create_distributed")
        self.device = device
        self.num_amplitudes_per_node = (1 << self.num_qubits) // self.num_nodes   # circuit.rs:135-136
        self.num_gpus_per_node = 1
        self.num_amplitudes_per_gpu = self.num_amplitudes_per_node

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle:
            self._lib.dvd_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state lifecycle (circuit.rs:262-322) ---------------------------------------------------
    def reset_amplitudes(self):
        _lib.check(self._lib.dvd_reset_zero_state(self._handle), "dvd_reset_zero_state")

    def reset(self):  # clears the gates but not `observables`, as the reference does (circuit.rs:303-306)
        self.reset_amplitudes()
        self.gates = []

    def set_parameters(self, parameters: Sequence[float]):
        idx = [i for i, g in enumerate(self.gates) if g[3] is not None]
        for i, p in zip(idx, parameters):
            self.gates[i][3] = float(p)

    # ---- gate list (circuit.rs:608-691) ----------------------------------------------------------
    def _check_qubit(self, q):
        if not (0 <= int(q) < self.num_qubits):
            raise ValueError(f"qubit {q} out of range for {self.num_qubits} qubits")
        return int(q)

    def add_hadamard_gate(self, active_qubit: int):
        self.gates.append(["Hadamard", self._check_qubit(active_qubit), None, None])

    def add_rotation_x_gate(self, active_qubit: int, theta: float):
        self.gates.append(["RotationX", self._check_qubit(active_qubit), None, float(theta)])

    def add_rotation_y_gate(self, active_qubit: int, theta: float):
        self.gates.append(["RotationY", self._check_qubit(active_qubit), None, float(theta)])

    def add_rotation_z_gate(self, active_qubit: int, theta: float):
        self.gates.append(["RotationZ", self._check_qubit(active_qubit), None, float(theta)])

    def _pauli(self, name, q, is_observable):
        self.gates.append([name, self._check_qubit(q), None, None])
        if is_observable:
            self.observables.append(len(self.gates) - 1)

    def add_pauli_x_gate(self, active_qubit: int, is_observable: bool):
        self._pauli("PauliX", active_qubit, is_observable)

    def add_pauli_y_gate(self, active_qubit: int, is_observable: bool):
        self._pauli("PauliY", active_qubit, is_observable)

    def add_pauli_z_gate(self, active_qubit: int, is_observable: bool):
        self._pauli("PauliZ", active_qubit, is_observable)

    def add_cnot_gate(self, control_qubit: int, target_qubit: int):
        c, t = self._check_qubit(control_qubit), self._check_qubit(target_qubit)
        if c == t:
            raise ValueError("control and target must differ")
        self.gates.append(["CNOT", t, c, None])

    # extensions over the reference's add_* set (gates.rs has S and T but no add_ method)
    def add_s_gate(self, active_qubit: int):
        self.gates.append(["S", self._check_qubit(active_qubit), None, None])

    def add_t_gate(self, active_qubit: int):
        self.gates.append(["T", self._check_qubit(active_qubit), None, None])

    # the FFI carries an arbitrary 2x2 and a control (circuit_gpu.rs:31-60); the reference's add_* set never exposes it
    def add_unitary_gate(self, active_qubit: int, matrix):
        """Arbitrary 2x2 (row-major, complex) on one qubit."""
        self.gates.append(["Unitary", self._check_qubit(active_qubit), None, None, gates.from_2x2(matrix)])

    def add_controlled_gate(self, control_qubit: int, target_qubit: int, matrix):
        """Arbitrary 2x2 on `target_qubit`, applied where `control_qubit` is 1 (the rule of circuit_multithreading.rs:36-38)."""
        c, t = self._check_qubit(control_qubit), self._check_qubit(target_qubit)
        if c == t:
            raise ValueError("control and target must differ")
        self.gates.append(["Unitary", t, c, None, gates.from_2x2(matrix)])

    def print_operations(self):  # circuit.rs:693-698 prints `operations`, which nothing ever fills
        pass

    # ---- forward (circuit.rs:341-375) ------------------------------------------------------------
    def _gate_array(self):
        obs = set(self.observables)
        todo = [g for i, g in enumerate(self.gates) if i not in obs]     # :347-349
        # `gates` / `observables` are public lists (the reference's fields): the marshalled array is reused only
        # while their contents are unchanged
        key = tuple(tuple(tuple(x) if isinstance(x, list) else x for x in g) for g in todo)
        if self._gate_cache is not None and self._gate_cache[0] == key:
            return self._gate_cache[1], len(todo)
        arr = (_lib.Gate * max(1, len(todo)))()
        for k, g in enumerate(todo):
            name, t, c, p = g[:4]
            arr[k].target = t
            arr[k].control = -1 if c is None else c
            arr[k].m[:] = gates.matrix(name, p, g[4] if len(g) > 4 else None)
        self._gate_cache = (key, arr)
        return arr, len(todo)

    def forward(self):
        """Apply every non-observable gate, in order, to the current state (no implicit reset)."""
        arr, n = self._gate_array()
        before = self.stats() if self.num_nodes > 1 else None
        _lib.check(self._lib.dvd_timer_begin(self._handle), "dvd_timer_begin")
        _lib.check(self._lib.dvd_apply_circuit(self._handle, arr, n), "dvd_apply_circuit")
        _lib.check(self._lib.dvd_flush(self._handle), "dvd_flush")
        ms = ctypes.c_double()
        _lib.check(self._lib.dvd_timer_end(self._handle, ctypes.byref(ms)), "dvd_timer_end")   # records + waits for the stream
        _lib.check(self._lib.dvd_synchronize(self._handle), "dvd_synchronize")
        self._profilers["forward"].add_ms(ms.value)
        if before is not None:
            after = self.stats()
            if after["global_swaps"] > before["global_swaps"]:
                self._profilers["inter_gpu"].add_ms((after["remap_ms"] + after["swap_ms"]) - (before["remap_ms"] + before["swap_ms"]))

    def forward_async(self):
        """forward() without the final device synchronisation (used by the bench harness)."""
        arr, n = self._gate_array()
        _lib.check(self._lib.dvd_apply_circuit(self._handle, arr, n), "dvd_apply_circuit")
        _lib.check(self._lib.dvd_flush(self._handle), "dvd_flush")
        return n

    def synchronize(self):
        _lib.check(self._lib.dvd_synchronize(self._handle), "dvd_synchronize")

    # ---- observation ----------------------------------------------------------------------------
    def state_numpy(self) -> np.ndarray:
        """Local chunk as a complex128 numpy array."""
        n = self.num_amplitudes_per_node
        re = np.empty(n, dtype=np.float64)
        im = np.empty(n, dtype=np.float64)
        _lib.check(self._lib.dvd_read_state(self._handle, _dp(re), _dp(im), 0, n), "dvd_read_state")
        return re + 1j * im

    def load_state_numpy(self, amplitudes, first: int = 0):
        """Overwrite local amplitudes [first, first + len) with a complex array (the reference's TODO'd
        load_amplitudes_local_on_device, rust_communication.cu:384-398, done properly)."""
        a = np.asarray(amplitudes, dtype=np.complex128)
        re = np.ascontiguousarray(a.real); im = np.ascontiguousarray(a.imag)
        _lib.check(self._lib.dvd_load_state(self._handle, _dp(re), _dp(im), int(first), a.shape[0]), "dvd_load_state")

    def retrieve_amplitudes_on_host(self):  # circuit.rs:380-405
        self._host_state = self.state_numpy()

    def get_real_part_state(self) -> List[float]:   # circuit.rs:596-598
        return self.state_numpy().real.tolist()

    def get_imaginary_part_state(self) -> List[float]:   # circuit.rs:600-602
        return self.state_numpy().imag.tolist()

    def measure_numpy(self) -> np.ndarray:
        n = self.num_amplitudes_per_node
        p = np.empty(n, dtype=np.float64)
        _lib.check(self._lib.dvd_probabilities(self._handle, _dp(p), 0, n), "dvd_probabilities")
        return p

    def measure(self) -> List[float]:   # circuit.rs:565-589: this node's probabilities
        return self.measure_numpy().tolist()

    def norm(self) -> float:
        out = ctypes.c_double()
        _lib.check(self._lib.dvd_norm(self._handle, ctypes.byref(out)), "dvd_norm")
        return out.value

    def set_sampler(self, order: str):
        """'tree' (default: pairwise summation tree, any size) or 'sequential' (the reference's strict left-to-right
        cumulative sums, utils.rs:270-274: the reference's indices for every draw; at most 30 local qubits)."""
        orders = {"tree": 0, "sequential": 1}
        if order not in orders:
            raise ValueError(f"sampler order must be one of {sorted(orders)}")
        _lib.check(self._lib.dvd_set_sampler(self._handle, orders[order]), "dvd_set_sampler")

    def sample_numpy(self, num_samples: Optional[int] = None, uniforms=None) -> np.ndarray:
        shots = 1000 if num_samples is None else int(num_samples)          # circuit.rs:439-443
        per_shot = 2 if self.num_nodes > 1 else 1
        if uniforms is None:
            u = np.random.default_rng().random(shots * per_shot) if self.rank == 0 else None
            if self.num_nodes > 1:
                u = distributed.broadcast_array(u, (shots * per_shot,), np.float64)
        else:
            u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
        if u.shape[0] < shots * per_shot:
            raise ValueError(f"need {shots * per_shot} uniforms for {shots} shots on {self.num_nodes} rank(s)")
        u = np.ascontiguousarray(u[: shots * per_shot])
        out = np.empty(shots, dtype=np.uint64)
        self._profilers["sampling"].start()
        _lib.check(self._lib.dvd_sample(self._handle, _dp(u), shots,
                                        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))), "dvd_sample")
        self._profilers["sampling"].stop()
        return out

    def sample(self, num_samples: Optional[int] = None, uniforms=None) -> List[int]:   # circuit.rs:434-458
        return [int(x) for x in self.sample_numpy(num_samples, uniforms)]

    def sample_local(self, num_samples: int, node_probabilities, uniforms=None) -> List[int]:
        """circuit.rs:468-485 + utils.rs:258-277 on a caller-supplied probability list (host only)."""
        p = np.asarray(node_probabilities, dtype=np.float64)
        cum = np.concatenate([[0.0], np.cumsum(p)])
        u = np.random.default_rng().random(num_samples) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        idx = np.searchsorted(cum, u[:num_samples] * cum[-1], side="left")
        return [int(max(i - 1, 0)) for i in idx]

    def extract_expectation_values_numpy(self, samples) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(samples, dtype=np.uint64))
        if s.shape[0] == 0:          # the reference's loop body never runs (circuit.rs:496)
            return np.empty((0, len(self.observables)), dtype=np.float64)
        q = np.ascontiguousarray(np.array([self.gates[i][1] for i in self.observables], dtype=np.int32))
        out = np.empty((s.shape[0], q.shape[0]), dtype=np.float64)
        _lib.check(self._lib.dvd_extract_expectation_values(
            self._handle, s.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), s.shape[0],
            q.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), q.shape[0], _dp(out)),
            "dvd_extract_expectation_values")
        return out

    def extract_expectation_values(self, samples) -> List[List[float]]:   # circuit.rs:494-513
        return self.extract_expectation_values_numpy(samples).tolist()

    def expectation_z(self) -> np.ndarray:
        """Exact <Z_q> for every qubit (extension, no sampling noise)."""
        out = np.empty(self.num_qubits, dtype=np.float64)
        _lib.check(self._lib.dvd_expectation_z(self._handle, _dp(out)), "dvd_expectation_z")
        return out

    def get_fidelity_between_two_states_with_parameters(self, parameters_1, parameters_2) -> float:
        """circuit.rs:753-769 + circuit_metrics.rs:12-92: |<psi(p1)|psi(p2)>|^2, then reset().  The first state stays on
        the device as a snapshot (same device / rank / communicator); on distributed states the partial dots are
        allreduced over the ranks (the reference's distributed_dot gathers them on the root and broadcasts)."""
        other = ctypes.c_void_p()
        self.reset_amplitudes(); self.set_parameters(parameters_1); self.forward()
        _lib.check(self._lib.dvd_snapshot(self._handle, ctypes.byref(other)), "dvd_snapshot")
        try:
            self.reset_amplitudes(); self.set_parameters(parameters_2); self.forward()
            out = ctypes.c_double()
            _lib.check(self._lib.dvd_fidelity(other, self._handle, ctypes.byref(out)), "dvd_fidelity")
        finally:
            self._lib.dvd_destroy(other)
        self.reset()                                                        # circuit_metrics.rs:30
        return out.value

    # ---- profiling getters (circuit.rs:700-751) ---------------------------------------------
    def get_profiling_results_forward(self):
        p = self._profilers["forward"]; return (p.iterations, p.mean())

    def get_profiling_results_inter_node_communications(self):
        p = self._profilers["inter_node"]; return (p.iterations, p.mean())

    def get_profiling_results_inter_gpu_communications(self):
        p = self._profilers["inter_gpu"]; return (p.iterations, p.mean())

    def get_profiling_results_sampling(self):
        p = self._profilers["sampling"]; return (p.iterations, p.mean())

    def print_profiling_results(self):
        print("profiling results")
        for label, key in (("Forward", "forward"), ("InterNodeCommunications", "inter_node"),
                           ("InterGPUCommunications", "inter_gpu"), ("Sampling", "sampling")):
            p = self._profilers[key]
            print(f"profiling {label}: iterations {p.iterations} elapsed time {p.mean()}")

    # ---- engine introspection ----------------------------------------------------------------
    def stats(self) -> dict:
        st = _lib.Stats()
        _lib.check(self._lib.dvd_get_stats(self._handle, ctypes.byref(st)), "dvd_get_stats")
        return st.as_dict()

    def stats_reset(self):
        _lib.check(self._lib.dvd_stats_reset(self._handle), "dvd_stats_reset")

    def set_jit(self, mode: int):
        """Structure-specialised pass kernels compiled at run time (csrc/jit.h): 0 off, 1 background, 2 on first use."""
        _lib.check(self._lib.dvd_set_jit(self._handle, int(mode)), "dvd_set_jit")

    def jit_wait(self):
        """Block until every queued kernel compilation has finished."""
        _lib.check(self._lib.dvd_jit_wait(self._handle), "dvd_jit_wait")

    def jit_info(self) -> dict:
        c, f, q = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        sec = ctypes.c_double()
        msg = ctypes.create_string_buffer(4096)
        _lib.check(self._lib.dvd_jit_info(self._handle, ctypes.byref(c), ctypes.byref(f), ctypes.byref(q), ctypes.byref(sec), msg, 4096),
                   "dvd_jit_info")
        forms = (ctypes.c_int64 * 7)()
        _lib.check(self._lib.dvd_jit_forms(forms), "dvd_jit_forms")
        names = ("classic2", "classic3", "ring")
        return {"compiled": c.value, "failed": f.value, "pending": q.value, "compile_seconds": sec.value,
                "message": msg.value.decode(errors="replace"), "tuning": int(forms[0]),
                "chosen": {n: int(forms[1 + i]) for i, n in enumerate(names)},
                "launches": {n: int(forms[4 + i]) for i, n in enumerate(names)}}

    def set_unfused(self, flag: bool):
        _lib.check(self._lib.dvd_set_unfused(self._handle, int(bool(flag))), "dvd_set_unfused")

    def timer_begin(self):
        _lib.check(self._lib.dvd_timer_begin(self._handle), "dvd_timer_begin")

    def timer_end(self) -> float:
        ms = ctypes.c_double()
        _lib.check(self._lib.dvd_timer_end(self._handle, ctypes.byref(ms)), "dvd_timer_end")
        return ms.value


def initialize_mpi():
    """src/lib.rs:16-19.  Here: make sure the rank-per-GPU process group exists."""
    distributed.initialize()
