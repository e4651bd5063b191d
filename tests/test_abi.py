"""The C-ABI library loads without a GPU, exports every declared symbol and fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

from damavand_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", txt)) - {"defined"})


def test_every_declared_symbol_is_exported():
    L = _lib.load()
    names = _declared("damavand_b200.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_reference_export_names_are_exported():
    L = _lib.load()
    names = [n for n in _declared("damavand_gpu_compat.h")]
    # the 14 names of /root/reference/damavand-gpu/rust_communication.cu
    for n in ["get_number_of_available_gpus", "get_memory_for_gpu", "peer_access_allowed", "print_timers",
              "exchange_amplitudes_between_gpus", "init_quantum_state", "sequential_measure_on_gpu",
              "concurrent_measure_on_gpu", "measure_on_gpu", "apply_one_qubit_gate_gpu_local",
              "apply_one_qubit_gate_gpu_distributed", "load_amplitudes_local_on_device",
              "split_amplitudes_between_gpus", "retrieve_amplitudes_on_host"]:
        assert n in names and hasattr(L, n), n


def test_no_silent_fallback_without_gpu():
    L = _lib.load()
    if L.dvd_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = L.dvd_create(4, 0, ctypes.byref(h))
    assert rc != 0 and not h
    assert b"GPU" in L.dvd_last_error() or len(L.dvd_last_error()) > 0
    from damavand_b200 import Circuit, DamavandError
    with pytest.raises(DamavandError):
        Circuit(4, "gpu")


def test_apply_method_strings():
    from damavand_b200 import Circuit
    for m in ("brute_force", "shuffle", "multithreading", "distributed_cpu", None):
        with pytest.raises(NotImplementedError):
            Circuit(3, m)
    with pytest.raises(ValueError, match="Apply method not recognized"):
        Circuit(3, "element_wise")       # the reference panics on this (circuit.rs:112,878)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "damavand_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the CPU oracle", "").replace("CPU oracle", ""), os.path.join(dirpath, f)


def test_planner_debug_entry_points():
    L = _lib.load()
    n = 14
    gates = (_lib.Gate * 3)()
    from damavand_b200 import gates as pg
    gates[0].target, gates[0].control = 13, -1; gates[0].m[:] = pg.hadamard()
    gates[1].target, gates[1].control = 13, 2; gates[1].m[:] = pg.pauli_x()
    gates[2].target, gates[2].control = 5, -1; gates[2].m[:] = pg.rotation_z(0.3)
    out = (ctypes.c_int32 * 256)()
    k = L.dvd_plan_debug(n, n, gates, 3, 0, out, 256)
    assert k > 0 and out[0] == 1              # one pass
    tile = list(out[1:13])
    assert 13 in tile and tile[:3] == [0, 1, 2]
    assert 3 <= out[14] <= 4                   # H, RZ phase (+ the pass constant), permuting switch
    # distributed planner: gate on a rank-index qubit forces a swap, identity restored afterwards
    perm = (ctypes.c_int32 * n)(*range(n))
    k = L.dvd_plan_distributed_debug(n, n - 1, gates, 3, perm, 1, out, 256)
    assert k > 0 and list(perm) == list(range(n))
    steps = out[0]
    assert steps >= 3
    # too-small buffer reports the needed size
    assert L.dvd_plan_debug(n, n, gates, 3, 0, out, 2) < 0
    # bad input is an error, not a crash
    gates[0].target = 99
    assert L.dvd_plan_debug(n, n, gates, 3, 1, out, 256) < -10**9


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the
    keys of the bench contract; a tiny workload keeps it at a second."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "layered14x2",
                          "--steps", "2", "--warmup", "1", "--cpu-gates-per-step", "20"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "gates_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64"


def test_jit_generator_and_nvrtc_compile():
    """The generated source of a pass's structure-specialised kernel compiles with NVRTC for sm_100a on a host
    without a GPU, and two passes with the same op structure (different angles) produce the same source."""
    import ctypes
    from damavand_b200 import _lib, circuits
    from oracle.oracle import OracleCircuit
    from tests.helpers import gate_array
    L = _lib.load()
    srcs = []
    for seed_shift in (0.0, 0.37):
        o = OracleCircuit.__new__(OracleCircuit); o.num_qubits = 14; o.gates = []; o.observables = []
        circuits.hea(o, 14, 2)
        for g in o.gates:
            if g.parameter is not None:
                g.parameter += seed_shift
        arr, ng = gate_array(o)
        buf = ctypes.create_string_buffer(1 << 20)
        k = L.dvd_jit_debug_source(14, 14, arr, ng, 0, 0, buf, 1 << 20)
        assert k > 0, L.dvd_last_error()
        srcs.append(buf.value)
    assert srcs[0] == srcs[1]
    assert b"apply_op<C_ALL>" in srcs[0] and b"tile_store<" in srcs[0]
    size = L.dvd_jit_debug_compile(srcs[0])
    assert size > 10000, L.dvd_last_error()
    assert L.dvd_jit_debug_compile(b"this is not CUDA") == -1 and b"error" in L.dvd_last_error()
    # the other two kernel forms of the same pass (3 CTAs per SM; two-group persistent ring) compile too
    k = L.dvd_jit_debug_source(14, 14, arr, ng, 0, 1, buf, 1 << 20)
    assert k > 0 and buf.value != srcs[0] and b"__launch_bounds__(NTHREADS, 3)" in buf.value
    assert L.dvd_jit_debug_compile(buf.value) > 10000, L.dvd_last_error()
    k = L.dvd_jit_debug_source(14, 14, arr, ng, 0, 2, buf, 1 << 20)
    assert k > 0 and buf.value != srcs[0] and b"ring_fetch(ring, slot + 3" in buf.value and b"group_sync(grp)" in buf.value
    assert L.dvd_jit_debug_compile(buf.value) > 10000, L.dvd_last_error()
    # the store-side-remap variant (the layout restore riding on the last pass of a distributed gate list): a
    # compile-time variant, so that the plain kernels do not pay registers for it
    assert b"true>(amp, pd" not in srcs[0]
    k = L.dvd_jit_debug_source(14, 14, arr, ng, 0, 16, buf, 1 << 20)
    assert k > 0 and buf.value != srcs[0] and b", true>(amp, pd, a, gbase - pd.rank_bits, s_toff)" in buf.value
    assert L.dvd_jit_debug_compile(buf.value) > 10000, L.dvd_last_error()


def test_jit_disk_cache(tmp_path, monkeypatch):
    """A compiled cubin is written to DVD_JIT_CACHE_DIR (atomically) and served from there the next time; a corrupt
    file is ignored and overwritten."""
    import time
    from damavand_b200 import _lib
    L = _lib.load()
    monkeypatch.setenv("DVD_JIT_CACHE_DIR", str(tmp_path / "jit"))
    src = b'extern "C" __global__ void dvd_pass_static(double* p) { p[threadIdx.x] *= 2.0; }\n// ' + str(time.time()).encode()
    t0 = time.perf_counter(); a = L.dvd_jit_debug_compile(src); t_cold = time.perf_counter() - t0
    files = list((tmp_path / "jit").glob("*.cubin"))
    assert a > 0 and len(files) == 1 and files[0].stat().st_size == a
    t0 = time.perf_counter(); b = L.dvd_jit_debug_compile(src); t_warm = time.perf_counter() - t0
    assert b == a and t_warm < t_cold
    files[0].write_bytes(b"garbage")
    assert L.dvd_jit_debug_compile(src) == a and files[0].stat().st_size == a
    monkeypatch.setenv("DVD_JIT_CACHE_DIR", "")
    assert L.dvd_jit_debug_compile(src + b" ") == a and len(list((tmp_path / "jit").glob("*"))) == 1


def test_rust_ffi_crate_declares_the_same_abi():
    """ffi/src/lib.rs cannot be compiled here (no cargo in the image): at least keep it in step with the header --
    every entry point of include/damavand_b200.h (the host-only planner / generator inspection calls aside) is declared,
    and dvd_stats has the same fields in the same order."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "damavand_b200.h")).read()
    rs = open(os.path.join(root, "ffi", "src", "lib.rs")).read()
    names = set(re.findall(r"^(?:int|double|int64_t|const char\*)\s+(dvd_\w+)\(", hdr, flags=re.M))
    assert len(names) > 35
    debug_only = {n for n in names if "debug" in n}
    missing = sorted(n for n in names - debug_only if f"pub fn {n}(" not in rs)
    assert missing == [], missing
    c_fields = re.findall(r"^\s+(?:int64_t|double)\s+(\w+);", hdr[hdr.index("typedef struct {"):hdr.index("} dvd_stats;")], flags=re.M)
    rs_fields = re.findall(r"pub (\w+): (?:i64|c_double),", rs[rs.index("pub struct dvd_stats"):rs.index("pub const DVD_SAMPLER_TREE")])
    assert c_fields == rs_fields and len(c_fields) >= 16
    from damavand_b200 import _lib
    assert [f for f, _ in _lib.Stats._fields_] == c_fields
