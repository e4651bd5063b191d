"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu.

Tolerances (BASELINE.json north_star): amplitudes and probabilities within 1e-12 relative,
sampled indices bit-exact under injected uniforms."""
import math

import numpy as np
import pytest

from damavand_b200 import circuits
from oracle import oracle
from oracle.oracle import OracleCircuit
from tests.helpers import Recorder, l2_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def gpu_circuit(n):
    from damavand_b200 import Circuit
    return Circuit(n, "gpu")


def both(n, build):
    rec = Recorder(); build(rec)
    g = rec.replay(gpu_circuit(n)); o = rec.replay(OracleCircuit(n))
    g.forward(); o.forward()
    return g, o


def test_reference_golden_vector():
    g, _ = both(2, lambda c: (c.add_hadamard_gate(0), c.add_hadamard_gate(1), c.add_cnot_gate(0, 1)))
    assert np.abs(g.state_numpy() - 0.5).max() < 1e-15
    assert g.get_real_part_state() == pytest.approx([0.5] * 4, abs=1e-15)
    assert g.get_imaginary_part_state() == [0.0] * 4


def test_known_answers():
    r = 1 / math.sqrt(2)
    g, _ = both(2, lambda c: (c.add_hadamard_gate(0), c.add_cnot_gate(0, 1)))
    assert np.allclose(g.state_numpy(), [r, 0, 0, r], atol=1e-15)
    g, _ = both(1, lambda c: c.add_rotation_x_gate(0, math.pi))
    assert np.allclose(g.state_numpy(), [0, -1j], atol=1e-15)
    for q in (0, 3, 11, 12, 17):
        g, _ = both(18, lambda c: c.add_pauli_x_gate(q, False))
        p = g.measure_numpy()
        assert p[1 << q] == 1.0 and p.sum() == 1.0


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11])
def test_small_states_simple_kernel(n):
    g, o = both(n, lambda c: circuits.random_circuit(c, n, 60, seed=n))
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL


@pytest.mark.parametrize("n", [12, 14])
def test_every_target_every_control(n):
    worst = 0.0
    for t in range(n):
        for ctl in [None] + [q for q in range(n) if q != t]:
            def build(c):
                for q in range(n):
                    c.add_rotation_y_gate(q, 0.3 + 0.11 * q); c.add_rotation_z_gate(q, 0.2 + 0.07 * q)
                if ctl is None:
                    c.add_rotation_x_gate(t, 1.234)
                else:
                    c.add_cnot_gate(ctl, t)
                c.add_rotation_z_gate((t + 1) % n, 0.5)
            g, o = both(n, build)
            e = rel_err(g.state_numpy(), o.amplitudes())
            worst = max(worst, e)
            assert e < TOL, (t, ctl, e)
            g.close()
    print("worst rel err", worst)


@pytest.mark.parametrize("n,gates,seed", [(12, 300, 1), (13, 400, 2), (16, 500, 3), (20, 400, 4), (22, 300, 5), (24, 200, 6)])
def test_random_circuits(n, gates, seed):
    g, o = both(n, lambda c: circuits.random_circuit(c, n, gates, seed))
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    assert l2_err(g.state_numpy(), o.amplitudes()) < TOL
    assert abs(g.norm() - 1.0) < 1e-12
    st = g.stats()
    assert st["gates_applied"] == gates and st["tile_passes"] < gates and st["kernel_launches"] > 0


def test_fused_equals_unfused():
    n = 16
    rec = Recorder(); circuits.random_circuit(rec, n, 300, 11)
    a = rec.replay(gpu_circuit(n)); b = rec.replay(gpu_circuit(n)); b.set_unfused(True)
    a.forward(); b.forward()
    assert rel_err(a.state_numpy(), b.state_numpy()) < TOL
    assert b.stats()["simple_passes"] == 300 and a.stats()["simple_passes"] == 0


def test_cfg1_layered20_full_pipeline():
    # BASELINE configs[0]: forward, measure, sample 1000, extract_expectation_values
    n = 20
    rec = Recorder(); circuits.layered(rec, n, 10)
    g = rec.replay(gpu_circuit(n)); o = rec.replay(OracleCircuit(n))
    g.forward(); o.forward()
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    pg, po = g.measure_numpy(), o.measure_np()
    assert np.abs(pg - po).max() <= TOL * po.max()
    u = np.random.default_rng(1235).random(1000)
    sg = g.sample(1000, uniforms=u)
    assert sg == o.sample(1000, uniforms=u, mode="tree")          # the specified summation order
    assert sg == o.sample(1000, uniforms=u, mode="sequential")    # and the reference's sequential order
    assert g.extract_expectation_values(sg) == o.extract_expectation_values(sg)


def test_shapes_qft_and_hea():
    g, o = both(18, lambda c: circuits.qft_like(c, 18))
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    g, o = both(18, lambda c: circuits.hea(c, 18, 8))
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL


@pytest.mark.parametrize("n", [3, 9, 10, 13, 21])
def test_sampler_bit_exact(n):
    g, o = both(n, lambda c: circuits.layered(c, n, 2, seed=n))
    u = np.random.default_rng(99).random(20000)
    u[0] = 0.0
    u[1] = np.nextafter(1.0, 0.0)
    sg = g.sample_numpy(u.size, u)
    p = o.measure_np()
    assert (sg == oracle.sample_tree(p, u)).all()
    assert (sg == oracle.sample_sequential(p, u)).all()


def test_sampler_on_exact_probabilities():
    # the sampler's own |amp|^2 must be the reference's norm_sqr bit for bit
    n = 15
    g, o = both(n, lambda c: circuits.random_circuit(c, n, 200, 21))
    ps = g.state_numpy()
    mine = ps.real * ps.real + ps.imag * ps.imag
    assert (g.measure_numpy() == mine).all()


def test_measure_and_sample_defaults():
    g, o = both(6, lambda c: [c.add_hadamard_gate(q) for q in range(6)])
    assert len(g.measure()) == 64 and abs(sum(g.measure()) - 1) < 1e-14
    s = g.sample()
    assert len(s) == 1000 and all(0 <= x < 64 for x in s)        # default 1000 shots (circuit.rs:439-443)
    assert g.sample(0) == []


def test_observables_forward_twice_reset_set_parameters():
    n = 12
    g = gpu_circuit(n); o = OracleCircuit(n)
    for c in (g, o):
        c.add_rotation_x_gate(0, 0.1); c.add_hadamard_gate(5); c.add_rotation_z_gate(5, 0.2)
        c.add_pauli_z_gate(0, True); c.add_cnot_gate(5, 11); c.add_pauli_x_gate(3, True)
        c.set_parameters([1.0, 2.0])
        c.forward(); c.forward()
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    assert g.extract_expectation_values([0, 1, 8, 9]) == o.extract_expectation_values([0, 1, 8, 9])
    g.reset_amplitudes()
    assert g.state_numpy()[0] == 1.0 and np.abs(g.state_numpy()[1:]).max() == 0.0
    g.reset()
    assert g.gates == [] and g.observables == [3, 5]
    assert g.extract_expectation_values([]) == []


def test_expectation_z_exact():
    n = 14
    g, o = both(n, lambda c: circuits.hea(c, n, 4))
    p = o.measure_np()
    idx = np.arange(p.size)
    want = np.array([(p * (1 - 2.0 * ((idx >> q) & 1))).sum() for q in range(n)])
    assert np.abs(g.expectation_z() - want).max() < 1e-12


def test_fidelity():
    n = 12
    g = gpu_circuit(n); o = OracleCircuit(n)
    for c in (g, o):
        for q in range(n):
            c.add_rotation_y_gate(q, 0.1)
        c.add_cnot_gate(0, 1)
    p1 = [0.1 * i for i in range(n)]; p2 = [0.1 * i + 0.05 for i in range(n)]
    assert abs(g.get_fidelity_between_two_states_with_parameters(p1, p2)
               - o.get_fidelity_between_two_states_with_parameters(p1, p2)) < 1e-12


def test_load_state_round_trip():
    from damavand_b200 import _lib
    import ctypes
    n = 13
    g = gpu_circuit(n)
    rng = np.random.default_rng(3)
    z = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    z /= np.linalg.norm(z)
    re, im = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    _lib.check(g._lib.dvd_load_state(g._handle, dp(re), dp(im), 0, 1 << n), "load")
    assert (g.state_numpy() == z).all()
    o = OracleCircuit(n); o.state = np.ascontiguousarray(z).view(np.float64).copy()
    for c in (g, o):
        c.add_hadamard_gate(12); c.add_cnot_gate(12, 0); c.forward()
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL


def test_error_paths():
    from damavand_b200 import DamavandError
    g = gpu_circuit(4)
    with pytest.raises(ValueError):
        g.add_hadamard_gate(4)
    with pytest.raises(ValueError):
        g.add_cnot_gate(1, 1)
    with pytest.raises(ValueError):
        g.sample(10, uniforms=[0.5])
    import ctypes
    re = (ctypes.c_double * 4)(1, 0, 0, 0); im = (ctypes.c_double * 4)()
    assert g._lib.dvd_apply_gate(g._handle, re, im, -1, 9) != 0
    assert b"target" in g._lib.dvd_last_error()
    with pytest.raises(DamavandError):
        from damavand_b200 import Circuit
        Circuit(41, "gpu")


def test_reference_export_names_drive_the_engine():
    """The reference's own FFI names (rust_communication.cu) on top of the engine."""
    import ctypes
    from damavand_b200 import _lib, gates as pg
    L = _lib.load()
    n = 13
    L.init_quantum_state.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    dp = ctypes.POINTER(ctypes.c_double)
    L.apply_one_qubit_gate_gpu_local.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.retrieve_amplitudes_on_host.argtypes = [ctypes.c_int, dp, dp]
    L.measure_on_gpu.argtypes = [ctypes.c_int, dp]
    assert L.get_number_of_available_gpus() == 1
    L.init_quantum_state(1 << n, 1, 1)
    o = OracleCircuit(n)
    for (name, t, c, p) in [("Hadamard", 12, None, None), ("CNOT", 3, 12, None), ("RotationY", 3, None, 0.4)]:
        m = pg.matrix(name, p)
        re = (ctypes.c_double * 4)(*m[0::2]); im = (ctypes.c_double * 4)(*m[1::2])
        L.apply_one_qubit_gate_gpu_local(re, im, n, 1 << n, -1 if c is None else c, t)
    o.add_hadamard_gate(12); o.add_cnot_gate(12, 3); o.add_rotation_y_gate(3, 0.4); o.forward()
    re = np.empty(1 << n); im = np.empty(1 << n); pr = np.empty(1 << n)
    L.retrieve_amplitudes_on_host(1 << n, re.ctypes.data_as(dp), im.ctypes.data_as(dp))
    L.measure_on_gpu(1 << n, pr.ctypes.data_as(dp))
    assert rel_err(re + 1j * im, o.amplitudes()) < TOL
    assert np.abs(pr - o.measure_np()).max() < 1e-15


@pytest.mark.parametrize("name,n", [("hea", 28), ("qft", 26)])
def test_large_state_properties(name, n):
    """Full-size style checks through size-independent properties: unitarity, <Z> consistency,
    fused == unfused on a subsample of amplitudes."""
    rec = Recorder()
    if name == "hea":
        circuits.hea(rec, n, 3)
    else:
        circuits.qft_like(rec, n)
    a = rec.replay(gpu_circuit(n)); a.forward()
    assert abs(a.norm() - 1.0) < 1e-11
    ez = a.expectation_z()
    assert np.all(np.abs(ez) <= 1 + 1e-12)
    b = rec.replay(gpu_circuit(n)); b.set_unfused(True); b.forward()
    import ctypes
    cnt = 1 << 16
    dp = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for first in (0, (1 << n) // 3, (1 << n) - cnt):
        ra, ia, rb, ib = (np.empty(cnt) for _ in range(4))
        a._lib.dvd_read_state(a._handle, dp(ra), dp(ia), first, cnt)
        b._lib.dvd_read_state(b._handle, dp(rb), dp(ib), first, cnt)
        scale = 2.0 ** (-n / 2)
        assert np.abs((ra - rb) + 1j * (ia - ib)).max() < 1e-12 * max(scale, np.abs(ra + 1j * ia).max())
    assert np.abs(ez - b.expectation_z()).max() < 1e-11


@pytest.fixture
def env(monkeypatch):
    """Engine knobs are read when a state is created (dvd_create -> kernels_init)."""
    def set_(**kw):
        for k, v in kw.items():
            monkeypatch.setenv(k, str(v))
    return set_


@pytest.mark.parametrize("kind,n,ring,lazy", [
    ("random", 21, 1, 1), ("random", 21, 1, 0), ("random", 21, 0, 1), ("random", 21, 0, 0),
    ("qft", 21, 1, 1), ("qft", 21, 0, 1), ("hea", 21, 1, 1), ("layered", 21, 1, 0), ("qft", 13, 1, 0), ("random", 12, 1, 0)])
def test_kernel_forms_agree_with_oracle(env, kind, n, ring, lazy):
    """Two-group persistent ("ring", cp.async prefetch through three shared-memory buffers) and one-tile-per-CTA forms
    of the interpreter pass kernel, with and without the lazy |0..0> input, against the oracle (2^21 amplitudes = 512
    tiles so that every ring really turns; 13 and 12 qubits = fewer tiles than ring slots)."""
    env(DVD_RING=ring, DVD_RING_MIN_TILES=1, DVD_LAZY_ZERO=lazy)
    build = {"random": lambda c: circuits.random_circuit(c, n, 120, 21),
             "qft": lambda c: circuits.qft_like(c, n),
             "hea": lambda c: circuits.hea(c, n, 2),
             "layered": lambda c: circuits.layered(c, n, 2)}[kind]
    g, o = both(n, build)
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    if kind == "random":   # a second forward starts from a non-zero state (no lazy input)
        g.forward(); o.forward()
        assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    g.reset_amplitudes(); g.forward()   # reset + forward reproduces the first result bit for bit
    first = o.amplitudes() if kind != "random" else None
    if first is not None:
        assert rel_err(g.state_numpy(), first) < TOL


def test_lazy_reset_observation_points(env):
    """reset_amplitudes() only marks the state as |0..0>; every observation must still see it."""
    import ctypes
    env(DVD_LAZY_ZERO=1)
    n = 14
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    g = gpu_circuit(n)
    p = g.measure_numpy()
    assert p[0] == 1.0 and p.sum() == 1.0
    assert g.norm() == 1.0
    assert g.sample(5, uniforms=[0.0, 0.3, 0.5, 0.9, 1.0 - 2 ** -53]) == [0] * 5
    assert list(g.expectation_z()) == [1.0] * n
    g.add_hadamard_gate(13); g.forward()
    g.reset_amplitudes()
    z = g.state_numpy()
    assert z[0] == 1.0 and np.abs(z[1:]).max() == 0.0
    g.reset_amplitudes()
    re, im = np.full(16, 0.25), np.zeros(16)
    from damavand_b200._lib import check as _check
    _check(g._lib.dvd_load_state(g._handle, dp(re), dp(im), 32, 16), "load")   # partial load after a lazy reset
    z = g.state_numpy()
    assert z[0] == 1.0 and (z[32:48] == 0.25).all() and np.abs(z[48:]).max() == 0.0 and np.abs(z[1:32]).max() == 0.0
    # two fresh states: <0|0> = 1
    a, b = gpu_circuit(n), gpu_circuit(n)
    f = ctypes.c_double()
    _check(a._lib.dvd_fidelity(a._handle, b._handle, ctypes.byref(f)), "fidelity")
    assert f.value == 1.0
    # unfused mode after a lazy reset goes through the one-gate kernel on a materialised state
    c = gpu_circuit(n); c.set_unfused(True); c.add_pauli_x_gate(7, False); c.forward()
    assert c.measure_numpy()[1 << 7] == 1.0


def test_plan_cache_reuses_only_identical_gate_lists():
    n = 14
    g = gpu_circuit(n); o = OracleCircuit(n)
    for c in (g, o):
        circuits.hea(c, n, 2)
    g.forward(); o.forward()
    assert g.stats()["plan_cache_hits"] == 0
    g.reset_amplitudes(); g.forward()                      # same gates: the plan and the marshalled array are reused
    assert g.stats()["plan_cache_hits"] == 1
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
    params = [0.01 * i for i in range(2 * n * 2)]
    for c in (g, o):
        c.set_parameters(params); c.reset_amplitudes(); c.forward()
    assert g.stats()["plan_cache_hits"] == 1               # new angles: planned again
    assert rel_err(g.state_numpy(), o.amplitudes()) < TOL


@pytest.mark.parametrize("n", [13, 20])
def test_support_tracking_partial_circuits(n):
    """After a reset only amplitude 0 is stored; passes launch only the tiles the touched qubits can populate and
    the implied zeros appear at the first observation.  Circuits that leave most qubits in |0>, observed in the
    middle and continued."""
    def diag_only(c):
        c.add_rotation_z_gate(3, 0.4); c.add_pauli_z_gate(n - 1, False)
    def x_moves(c):
        c.add_pauli_x_gate(n - 1, False); c.add_pauli_x_gate(0, False); c.add_cnot_gate(n - 1, 7); c.add_cnot_gate(3, 8)
    def one_high(c):
        c.add_hadamard_gate(n - 2); c.add_cnot_gate(n - 2, 2); c.add_cnot_gate(9, 5); c.add_rotation_y_gate(n - 3, 0.3)
    def growing(c):
        for q in np.random.default_rng(n).permutation(n)[:9]:
            c.add_rotation_x_gate(int(q), 0.1 + 0.05 * q); c.add_rotation_z_gate(int(q), 0.2)
            c.add_cnot_gate(int(q), int((q + 5) % n))
    def ghz(c):
        c.add_hadamard_gate(n - 1)
        for q in range(n - 1, 0, -1):
            c.add_cnot_gate(q, q - 1)
    for build in (diag_only, x_moves, one_high, growing, ghz):
        g, o = both(n, build)
        assert g.norm() == pytest.approx(1.0, abs=1e-13)                 # observation in the middle
        assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
        for c in (g, o):                                                # continue on the (now dense) state
            c.gates = []
            c.add_hadamard_gate(1); c.add_cnot_gate(1, n - 1); c.add_rotation_x_gate(n // 2, 0.7)
            c.forward()
        assert rel_err(g.state_numpy(), o.amplitudes()) < TOL
        u = np.random.default_rng(3).random(64)
        assert g.sample(64, uniforms=u) == o.sample(64, uniforms=u, mode="tree")


def _require_jit(info):
    """Run-time compilation must work wherever the CUDA toolkit ships NVRTC: a silent skip would let the specialised
    kernels rot.  Only a machine with no libnvrtc at all skips."""
    if not info["message"]:
        return
    import glob
    have = glob.glob("/usr/local/cuda*/lib64/libnvrtc.so*") + glob.glob("/usr/local/cuda*/targets/*/lib/libnvrtc.so*")
    try:
        import nvidia.cuda_nvrtc as m   # the pip wheel torch depends on
        import os
        have += glob.glob(os.path.join(os.path.dirname(m.__file__), "lib", "libnvrtc.so*"))
    except Exception:
        pass
    if have:
        pytest.fail(f"NVRTC is installed ({have[0]}) but the engine cannot use it: {info['message']}")
    pytest.skip("run-time compilation unavailable (no libnvrtc on this machine): " + info["message"])


@pytest.mark.parametrize("form", ["classic2", "classic3", "ring"])
@pytest.mark.parametrize("kind,n", [("qft", 20), ("hea", 21), ("random", 21), ("qft", 13)])
def test_jit_kernel_forms_agree_with_oracle(env, kind, n, form):
    """Every kernel form the run-time generator can emit (one tile per CTA at 2 or 3 CTAs per SM, two-group persistent
    ring), forced one at a time, against the oracle -- from a reset (support-tracked passes run the classic form) and
    on the dense state a second forward starts from (the ring form proper)."""
    env(DVD_JIT_MIN_QUBITS=12, DVD_JIT_FORM=form, DVD_RING_MIN_TILES=1)
    build = {"qft": lambda c: circuits.qft_like(c, n), "hea": lambda c: circuits.hea(c, n, 2),
             "random": lambda c: circuits.random_circuit(c, n, 100, 77)}[kind]
    rec = Recorder(); build(rec)
    j = rec.replay(gpu_circuit(n)); o = rec.replay(OracleCircuit(n))
    j.set_jit(2)
    _require_jit(j.jit_info())
    before = j.jit_info()["launches"]
    j.forward(); o.forward()
    assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    j.forward(); o.forward()      # dense input
    assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    st, info = j.stats(), j.jit_info()
    assert st["jit_launches"] == st["tile_passes"] > 0, info
    assert info["failed"] == 0, info
    if form != "classic3":        # (a classic3 kernel that spills too much is dropped in favour of classic2)
        assert info["launches"][form] > before[form], info


def test_jit_form_is_chosen_by_measurement(env):
    """Without DVD_JIT_FORM the engine times its candidates on the first dense launches of a structure and keeps the
    fastest; results stay within tolerance of the oracle whichever form each launch used."""
    env(DVD_JIT_MIN_QUBITS=12, DVD_RING_MIN_TILES=1)
    n = 22
    rec = Recorder(); circuits.hea(rec, n, 3)
    j = rec.replay(gpu_circuit(n)); o = rec.replay(OracleCircuit(n))
    j.set_jit(2)
    _require_jit(j.jit_info())
    for _ in range(8):
        j.forward(); o.forward()
        j.synchronize()
    assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    info = j.jit_info()
    assert info["failed"] == 0 and info["tuning"] == 0 and sum(info["chosen"].values()) > 0, info


@pytest.mark.parametrize("kind,n", [("qft", 20), ("hea", 20), ("random", 21), ("partial", 20)])
def test_jit_specialised_kernels_agree_with_oracle(env, kind, n):
    """Structure-specialised pass kernels (NVRTC, compiled on first use here) against the oracle, and bit for bit
    against the interpreter kernels they replace; a second parameter set reuses the compiled kernels."""
    env(DVD_JIT_MIN_QUBITS=12)
    def build(c):
        if kind == "qft":
            circuits.qft_like(c, n)
        elif kind == "hea":
            circuits.hea(c, n, 2)
        elif kind == "random":
            circuits.random_circuit(c, n, 80, 31)
        else:
            c.add_hadamard_gate(n - 1); c.add_rotation_y_gate(3, 0.4); c.add_cnot_gate(n - 1, 5); c.add_rotation_z_gate(5, 0.3)
            c.add_rotation_x_gate(n - 2, 1.1); c.add_cnot_gate(3, n - 2)
    rec = Recorder(); build(rec)
    j = rec.replay(gpu_circuit(n)); i = rec.replay(gpu_circuit(n)); o = rec.replay(OracleCircuit(n))
    j.set_jit(2)
    info0 = j.jit_info()
    _require_jit(info0)
    j.forward(); i.forward(); o.forward()
    st = j.stats()
    assert st["jit_launches"] == st["tile_passes"] > 0, j.jit_info()
    assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    assert rel_err(j.state_numpy(), i.state_numpy()) < 1e-14
    # dense input, same structure: no new compilation (the count only grows when this process had not met the
    # structure before: an earlier test may have compiled it)
    compiled = j.jit_info()["compiled"]
    assert compiled >= info0["compiled"]
    j.forward(); o.forward()
    assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    if kind == "hea":     # new angles, same structure
        p = [0.05 * (k + 1) for k in range(2 * n * 2)]     # no zero angle: RY(0) is a different gate class
        for c in (j, o):
            c.set_parameters(p); c.reset_amplitudes(); c.forward()
        assert j.jit_info()["compiled"] <= compiled + 1        # (an angle-dependent gate class may add one)
        assert j.stats()["jit_launches"] == j.stats()["tile_passes"]
        assert rel_err(j.state_numpy(), o.amplitudes()) < TOL
    assert j.jit_info()["failed"] == 0


def _golden():
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join("tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_cuda_path_against_hand_derived_known_answers():
    mg = _golden()
    ka = np.load("tests/golden/known_answers.npz")
    for name in ka.files:
        g = mg.build_known(name, gpu_circuit)
        g.forward()
        assert np.abs(g.state_numpy() - ka[name]).max() < 1e-15, name


@pytest.mark.parametrize("kind", ["layered", "hea", "qft", "random"])
def test_cuda_path_against_committed_fixtures(kind):
    """12 qubits = the smallest state the tiled kernel runs on: amplitudes within 1e-12, sampled indices and
    expectation values bit-exact against tests/golden/oracle_regression.npz."""
    mg = _golden()
    fx = np.load("tests/golden/oracle_regression.npz")
    g = mg.regression_case(kind, gpu_circuit)
    g.forward()
    assert g.stats()["tile_passes"] > 0
    assert rel_err(g.state_numpy(), fx[f"{kind}_amplitudes"]) < TOL
    s = g.sample(64, uniforms=fx[f"{kind}_uniforms"])
    assert (np.asarray(s, dtype=np.uint64) == fx[f"{kind}_samples"]).all()
    assert (np.asarray(g.extract_expectation_values(s)) == fx[f"{kind}_expectation"]).all()


def _random_unitary(rng):
    z = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


@pytest.mark.parametrize("n", [10, 14, 20])
def test_unitary_and_controlled_gates(n):
    """add_unitary_gate / add_controlled_gate (arbitrary 2x2, optional control: what the reference's FFI carries,
    circuit_gpu.rs:31-60) on every kernel path: one-gate kernel (n < 12), tile kernel with register-, thread- and
    CTA-level controls."""
    rng = np.random.default_rng(n)
    def build(c):
        for q in range(n):
            c.add_hadamard_gate(q)
        for k in range(60):
            t = int(rng.integers(0, n)); ctl = int(rng.integers(0, n - 1)); ctl += ctl >= t
            u = _random_unitary(rng)
            if k % 3 == 0:
                c.add_unitary_gate(t, u.tolist())
            elif k % 3 == 1:
                c.add_controlled_gate(ctl, t, u.tolist())
            else:      # controlled diagonal (a true controlled phase) and a controlled non-unitary 2x2: the rule is linear algebra
                c.add_controlled_gate(ctl, t, [[1.0, 0.0], [0.0, np.exp(1j * rng.random())]] if k % 2 else (0.5 * u + 0.1).tolist())
    g, o = both(n, build)
    a, b = g.state_numpy(), o.amplitudes()
    assert rel_err(a, b) < TOL
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < TOL


def test_profiling_buckets_are_fed_with_device_time():
    """profiler.rs:3-9 / circuit.rs:700-751: (iterations, mean elapsed ns) per bucket; Forward carries CUDA-event time."""
    n = 22
    g = gpu_circuit(n)
    circuits.hea(g, n, 4)
    for _ in range(3):
        g.reset_amplitudes(); g.forward()
    it, mean_ns = g.get_profiling_results_forward()
    assert it == 3 and 1e4 < mean_ns < 1e9           # tens of microseconds .. well under a second
    g.sample(100, uniforms=np.linspace(0.0, 0.99, 100))
    it_s, mean_s = g.get_profiling_results_sampling()
    assert it_s == 1 and mean_s > 0
    assert g.get_profiling_results_inter_gpu_communications() == (0, 0.0)     # one GPU: nothing crosses NVLink
    assert g.get_profiling_results_inter_node_communications() == (0, 0.0)
    g.print_profiling_results()


def test_fidelity_between_parameter_sets():
    """circuit.rs:753-769: the first state stays resident as a snapshot; identical parameters give 1, and the value
    matches the oracle's two forwards."""
    n = 16
    rec = Recorder(); circuits.hea(rec, n, 2)
    g = rec.replay(gpu_circuit(n))
    n_par = sum(1 for c in rec.calls if "rotation" in c[0])
    p1 = [0.1 + 0.01 * k for k in range(n_par)]
    p2 = [0.2 + 0.013 * k for k in range(n_par)]
    assert abs(g.get_fidelity_between_two_states_with_parameters(p1, p1) - 1.0) < 1e-12
    g = rec.replay(gpu_circuit(n))
    f = g.get_fidelity_between_two_states_with_parameters(p1, p2)
    a = rec.replay(OracleCircuit(n)); a.set_parameters(p1); a.forward()
    b = rec.replay(OracleCircuit(n)); b.set_parameters(p2); b.forward()
    want = abs(np.vdot(a.amplitudes(), b.amplitudes())) ** 2
    assert abs(f - want) < 1e-12
    assert g.gates == []                                  # the reference resets the circuit afterwards (circuit_metrics.rs:30)


def test_expectation_z_beyond_32_local_qubits():
    """<Z_q> for local qubits >= 32 (a 33-qubit state is 128 GiB: fits one B200).  H on the top qubit, X on qubit 31,
    RY(pi/3) on qubit 5: <Z> = 0, -1, cos(pi/3), +1 elsewhere."""
    import ctypes
    from damavand_b200 import _lib
    L = _lib.load()
    if L.dvd_device_mem_mib(0) < 150 * 1024:
        pytest.skip("needs a device with > 150 GiB")
    n = 33
    try:
        g = gpu_circuit(n)
    except _lib.DamavandError as e:
        pytest.skip(f"cannot allocate 128 GiB here: {e}")
    g.add_hadamard_gate(n - 1); g.add_pauli_x_gate(31, False); g.add_rotation_y_gate(5, math.pi / 3)
    g.forward()
    ez = g.expectation_z()
    want = np.ones(n); want[n - 1] = 0.0; want[31] = -1.0; want[5] = math.cos(math.pi / 3)
    assert np.abs(ez - want).max() < 1e-12
    assert abs(g.norm() - 1.0) < 1e-12
    g.close()


@pytest.mark.parametrize("kind", ["random", "hea", "qft"])
def test_config_scale_parity_26_qubits(kind):
    """26 qubits (1 GiB state, the largest the oracle finishes in minutes): every amplitude against the oracle, in the
    max-relative and the l2-relative metric; then 10^5 shots on identical amplitudes (the oracle's, loaded into the
    engine) in both summation orders -- the default pairwise tree against the oracle's tree, the reference-order mode
    against the oracle's sequential cumulative sums (utils.rs:258-277), bit for bit -- and the two orders against each
    other: a shot may only differ where xsi falls within rounding of a bin edge, by one populated bin."""
    n = 26
    build = {"qft": lambda c: circuits.qft_like(c, n),                    # 1655 gates, all three passes' worth of twiddles
             "hea": lambda c: circuits.hea(c, n, 5),
             "random": lambda c: circuits.random_circuit(c, n, 200, 26)}[kind]
    g, o = both(n, build)
    a, b = g.state_numpy(), o.amplitudes()
    assert rel_err(a, b) < TOL
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < TOL
    assert np.abs(np.abs(a) ** 2 - np.abs(b) ** 2).max() / (np.abs(b) ** 2).max() < TOL      # probabilities
    g.load_state_numpy(b)
    p = o.measure_np()
    shots = 100000
    u = np.random.default_rng(2626).random(shots)
    u[:3] = [0.0, 1.0 - 2.0 ** -53, 0.5]
    s_tree = g.sample_numpy(shots, u)
    assert (s_tree == oracle.sample_tree(p, u)).all()
    g.set_sampler("sequential")
    s_seq = g.sample_numpy(shots, u)
    assert (s_seq == oracle.sample_sequential(p, u)).all()
    diff = np.nonzero(s_tree != s_seq)[0]
    assert diff.size <= 5, diff.size          # expected ~shots * |tree - sequential prefix| / bin width ~ 1e-3 at 26 qubits
    cum = np.cumsum(p)
    for k in diff:      # both answers bracket xsi to within the summation error of the prefix
        xsi = u[k] * cum[-1]
        for idx in (int(s_tree[k]), int(s_seq[k])):
            lo = cum[idx - 1] if idx else 0.0
            assert lo - 1e-12 <= xsi <= cum[idx] + 1e-12
    g.close()
