"""ctypes binding of libdamavand_b200.so (the C ABI in include/damavand_b200.h).

This is the stub a maintainer of the reference would replace with the Rust `extern "C"` block
(see INTEGRATION.md).  There is NO fallback: if the CUDA library is missing or a call fails, an
exception is raised.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DVD_LIB_PATH: development override used to A/B two builds of the same C ABI on one GPU box
LIB_PATH = os.environ.get("DVD_LIB_PATH") or os.path.join(HERE, "libdamavand_b200.so")

NCCL_ID_BYTES = 128


class DamavandError(RuntimeError):
    """Raised when a C-ABI call fails (the reference panics / exits here)."""


class Gate(ctypes.Structure):
    _fields_ = [("target", ctypes.c_int32), ("control", ctypes.c_int32), ("m", ctypes.c_double * 8)]


class Stats(ctypes.Structure):
    _fields_ = [
        ("gates_applied", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("tile_passes", ctypes.c_int64),
        ("simple_passes", ctypes.c_int64),
        ("stage_switches", ctypes.c_int64),
        ("global_swaps", ctypes.c_int64),
        ("swap_bytes_sent", ctypes.c_int64),
        ("pass_bytes", ctypes.c_double),
        ("gate_algorithmic_bytes", ctypes.c_double),
        ("plan_cache_hits", ctypes.c_int64),
        ("jit_launches", ctypes.c_int64),
        ("remap_passes", ctypes.c_int64),
        ("remap_bytes_in", ctypes.c_double),
        ("remap_ms", ctypes.c_double),
        ("swap_ms", ctypes.c_double),
        ("pass_fp64_instr", ctypes.c_double),
        ("store_remap_passes", ctypes.c_int64),
        ("store_remap_ms", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

_VP = ctypes.c_void_p
_DP = ctypes.POINTER(ctypes.c_double)
_I32P = ctypes.POINTER(ctypes.c_int32)
_U64P = ctypes.POINTER(ctypes.c_uint64)

# name -> (restype, argtypes); every symbol include/damavand_b200.h declares
SIGNATURES = {
    "dvd_device_count": (ctypes.c_int, []),
    "dvd_device_mem_mib": (ctypes.c_double, [ctypes.c_int]),
    "dvd_peer_access_allowed": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "dvd_last_error": (ctypes.c_char_p, []),
    "dvd_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_VP)]),
    "dvd_create_distributed": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, ctypes.POINTER(_VP)]),
    "dvd_nccl_unique_id": (ctypes.c_int, [_VP]),
    "dvd_destroy": (ctypes.c_int, [_VP]),
    "dvd_reset_zero_state": (ctypes.c_int, [_VP]),
    "dvd_apply_gate": (ctypes.c_int, [_VP, _DP, _DP, ctypes.c_int, ctypes.c_int]),
    "dvd_apply_circuit": (ctypes.c_int, [_VP, ctypes.POINTER(Gate), ctypes.c_int64]),
    "dvd_flush": (ctypes.c_int, [_VP]),
    "dvd_synchronize": (ctypes.c_int, [_VP]),
    "dvd_probabilities": (ctypes.c_int, [_VP, _DP, ctypes.c_int64, ctypes.c_int64]),
    "dvd_norm": (ctypes.c_int, [_VP, _DP]),
    "dvd_sample": (ctypes.c_int, [_VP, _DP, ctypes.c_int64, _U64P]),
    "dvd_set_sampler": (ctypes.c_int, [_VP, ctypes.c_int]),
    "dvd_extract_expectation_values": (ctypes.c_int, [_VP, _U64P, ctypes.c_int64, _I32P, ctypes.c_int32, _DP]),
    "dvd_expectation_z": (ctypes.c_int, [_VP, _DP]),
    "dvd_read_state": (ctypes.c_int, [_VP, _DP, _DP, ctypes.c_int64, ctypes.c_int64]),
    "dvd_load_state": (ctypes.c_int, [_VP, _DP, _DP, ctypes.c_int64, ctypes.c_int64]),
    "dvd_fidelity": (ctypes.c_int, [_VP, _VP, _DP]),
    "dvd_copy_state": (ctypes.c_int, [_VP, _VP]),
    "dvd_snapshot": (ctypes.c_int, [_VP, ctypes.POINTER(_VP)]),
    "dvd_num_qubits": (ctypes.c_int, [_VP]),
    "dvd_num_local_qubits": (ctypes.c_int, [_VP]),
    "dvd_rank": (ctypes.c_int, [_VP]),
    "dvd_world": (ctypes.c_int, [_VP]),
    "dvd_device": (ctypes.c_int, [_VP]),
    "dvd_get_stats": (ctypes.c_int, [_VP, ctypes.POINTER(Stats)]),
    "dvd_stats_reset": (ctypes.c_int, [_VP]),
    "dvd_timer_begin": (ctypes.c_int, [_VP]),
    "dvd_timer_end": (ctypes.c_int, [_VP, _DP]),
    "dvd_set_unfused": (ctypes.c_int, [_VP, ctypes.c_int]),
    "dvd_plan_debug": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Gate), ctypes.c_int64, ctypes.c_int, _I32P, ctypes.c_int64]),
    "dvd_set_jit": (ctypes.c_int, [_VP, ctypes.c_int]),
    "dvd_jit_wait": (ctypes.c_int, [_VP]),
    "dvd_jit_info": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), _DP, ctypes.c_char_p, ctypes.c_int64]),
    "dvd_jit_forms": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int64)]),
    "dvd_jit_debug_compile": (ctypes.c_int64, [ctypes.c_char_p]),
    "dvd_jit_debug_source": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Gate), ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int64]),
    "dvd_plan_distributed_debug": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(Gate), ctypes.c_int64, _I32P, ctypes.c_int, _I32P, ctypes.c_int64]),
}

# the reference's own export names (include/damavand_gpu_compat.h)
COMPAT_SYMBOLS = [
    "get_number_of_available_gpus", "get_memory_for_gpu", "peer_access_allowed", "print_timers",
    "exchange_amplitudes_between_gpus", "init_quantum_state", "sequential_measure_on_gpu",
    "concurrent_measure_on_gpu", "measure_on_gpu", "apply_one_qubit_gate_gpu_local",
    "apply_one_qubit_gate_gpu_distributed", "load_amplitudes_local_on_device",
    "split_amplitudes_between_gpus", "retrieve_amplitudes_on_host",
    "dvd_compat_set_distributed", "dvd_compat_state",
]


def load():
    """Load the CUDA library; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DamavandError(
            f"{LIB_PATH} is missing: build it with `python -m damavand_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for apply_method='gpu'/'distributed_gpu'."
        )
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dvd_last_error()
        raise DamavandError(f"{what}: {msg.decode() if msg else 'error ' + str(rc)}")
