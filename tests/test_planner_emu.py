"""Planner + tile-kernel logic, replayed on the CPU (tests/emu) against the oracle.

No GPU here: the replay includes the same tile_core.cuh the CUDA kernel is compiled from and the
real planner, so a failure means the CUDA path would be wrong too."""
import numpy as np
import pytest

from damavand_b200 import circuits
from oracle.oracle import OracleCircuit
from tests.helpers import emu, emu_run, rel_err

TOL = 1e-12


def _check(circ: OracleCircuit, world=1):
    got, stats = emu_run(circ, world)                    # with the diagonal-run fusion pre-pass
    got_raw, stats_raw = emu_run(circ, world, fuse=False)  # gate by gate
    got_sp, _ = emu_run(circ, world, track_support=True)   # from a reset: unwritten memory is NaN and never read
    circ.forward()
    assert rel_err(got, circ.amplitudes()) < TOL
    assert rel_err(got_raw, circ.amplitudes()) < TOL
    assert not np.isnan(got_sp.view(np.float64)).any()
    assert rel_err(got_sp, circ.amplitudes()) < TOL
    if world > 1:
        # the same plan with every global<->local swap as an exchange of its own (the fallback when the second
        # chunk does not fit): the default above lets the swaps ride on the next pass's load
        got_x, stats_x = emu_run(circ, world, fused_remap=False)
        got_xs, _ = emu_run(circ, world, track_support=True, fused_remap=False)
        assert rel_err(got_x, circ.amplitudes()) < TOL and rel_err(got_xs, circ.amplitudes()) < TOL
        assert stats_x["swaps"] == stats["swaps"]
    return stats


def test_bank_conflict_free_layouts():
    assert emu().emu_max_bank_conflict() == 1


def test_golden_vector_small_state():
    c = OracleCircuit(2); c.add_hadamard_gate(0); c.add_hadamard_gate(1); c.add_cnot_gate(0, 1)
    got, _ = emu_run(c)
    assert np.abs(got - 0.5).max() < 1e-15


@pytest.mark.parametrize("n", [12, 14])
def test_every_target_every_control(n):
    # a dense prefix so that every amplitude is non-zero and distinct, then ONE gate under test
    for t in range(n):
        for c in [None] + [q for q in range(n) if q != t]:
            circ = OracleCircuit(n)
            for q in range(n):
                circ.add_rotation_y_gate(q, 0.3 + 0.11 * q); circ.add_rotation_z_gate(q, 0.2 + 0.07 * q)
            if c is None:
                circ.add_rotation_x_gate(t, 1.234)
            else:
                circ.add_cnot_gate(c, t)
            circ.add_rotation_z_gate((t + 1) % n, 0.5)      # a diagonal gate after it
            _check(circ)


@pytest.mark.parametrize("kind", ["H", "X", "Y", "Z", "RX", "RY", "RZ", "S", "T"])
def test_every_gate_kind_every_position(kind):
    n = 13
    for t in range(n):
        circ = OracleCircuit(n)
        for q in range(n):
            circ.add_rotation_y_gate(q, 0.4 + 0.1 * q)
        add = {"H": lambda: circ.add_hadamard_gate(t), "X": lambda: circ.add_pauli_x_gate(t, False),
               "Y": lambda: circ.add_pauli_y_gate(t, False), "Z": lambda: circ.add_pauli_z_gate(t, False),
               "RX": lambda: circ.add_rotation_x_gate(t, 0.77), "RY": lambda: circ.add_rotation_y_gate(t, 0.77),
               "RZ": lambda: circ.add_rotation_z_gate(t, 0.77),
               "S": lambda: circ.gates.append(__import__("oracle.oracle", fromlist=["_Gate"])._Gate("S", t)),
               "T": lambda: circ.gates.append(__import__("oracle.oracle", fromlist=["_Gate"])._Gate("T", t))}[kind]
        add()
        _check(circ)


@pytest.mark.parametrize("n,gates,seed", [(12, 300, 1), (13, 400, 2), (15, 500, 3), (16, 300, 4)])
def test_random_circuits(n, gates, seed):
    circ = OracleCircuit(n)
    circuits.random_circuit(circ, n, gates, seed)
    stats = _check(circ)
    assert stats["ops"] <= gates + 4 * stats["passes"] and stats["passes"] < gates


def test_workload_shapes_small():
    c = OracleCircuit(13); g = circuits.qft_like(c, 13); s = _check(c)
    assert s["ops"] < g / 2 and s["passes"] <= 4      # CNOT-conjugated RZ runs fuse into parity phases
    c = OracleCircuit(14); g = circuits.hea(c, 14, 6); s = _check(c)
    assert s["ops"] <= g + 4 * s["passes"]      # diagonal gates are re-emitted as phase-polynomial ops
    c = OracleCircuit(12); g = circuits.layered(c, 12, 3); s = _check(c)
    assert s["ops"] <= g + 4 * s["passes"]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n", [15, 16])
def test_distributed_replay(world, n):
    circ = OracleCircuit(n)
    circuits.random_circuit(circ, n, 200, seed=10 + world)
    stats = _check(circ, world)
    assert stats["swaps"] >= 1          # gates on rank-index qubits force exchanges
    circ = OracleCircuit(n); circuits.qft_like(circ, n); _check(circ, world)
    circ = OracleCircuit(n); circuits.hea(circ, n, 3); _check(circ, world)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tail_deferral_and_store_side_restore(world):
    """The two schedule transformations of round 2, forced on and off: gates of a nearly empty last pass wait for the next
    layout (every threshold the planner tries), and swap rounds ride on the STORE of the pass in front of them
    (PassDesc::remap_st): only the layout restore (mode 1, the default) or every round that can (mode 2) -- against the
    oracle, dense and from a reset (NaN-filled memory)."""
    seen_defer = seen_store = seen_push = False
    for n, kind, gates, seed in ((16, "random", 200, 10 + world), (17, "random", 200, 10 + world), (18, "random", 300, 1),
                                 (16, "hea", 0, 0), (17, "qft", 0, 0)):
        circ = OracleCircuit(n)
        if kind == "random":
            circuits.random_circuit(circ, n, gates, seed=seed)
        elif kind == "hea":
            circuits.hea(circ, n, 3)
        else:
            circuits.qft_like(circ, n)
        ref = OracleCircuit(n); ref.gates = list(circ.gates); ref.forward()
        want = ref.amplitudes()
        base = None
        for defer in (-1, 0, 12, 32):
            for store in (1, 0, 2):
                got, st = emu_run(circ, world, store_side=store, defer=defer)
                assert rel_err(got, want) < TOL, (n, kind, defer, store)
                assert st["store_side"] <= (0, 1, 99)[store]
                if defer == 0 and not store:
                    base = st["passes"]
                if defer == -1:
                    seen_defer |= st["defer"] > 0
                    seen_store |= store == 1 and st["store_side"] == 1
                    seen_push |= store == 2 and st["store_side"] >= 2
                if defer in (-1, 12):
                    got, _ = emu_run(circ, world, store_side=store, defer=defer, track_support=True)
                    assert not np.isnan(got.view(np.float64)).any() and rel_err(got, want) < TOL, (n, kind, defer, store, "sparse")
        tuned = emu_run(circ, world)[1]
        assert tuned["passes"] <= base          # the tuned schedule never needs more passes than the plain one
    # (coverage of the transformations themselves; with victims going home inside the swap rounds many schedules end in
    # the reference layout and have no restore left to ride on a store)
    assert seen_defer and seen_push and (seen_store or world == 4)


@pytest.mark.parametrize("kind,n,arg,world", [("random", 19, 640, 8), ("hea", 19, 10, 8)])
def test_config_twins_of_the_multi_gpu_workloads(kind, n, arg, world):
    """Twins of BASELINE configs 4 and 5 (the same generators, gate counts and rank counts at 19 qubits): the schedule
    the engine derives for them -- deferred tails, several swap rounds on stores, the restore with local transpositions --
    replayed on emulated ranks, dense and from a reset."""
    c = OracleCircuit(n)
    if kind == "random":
        circuits.random_circuit(c, n, arg)
    else:
        circuits.hea(c, n, arg)
    got, st = emu_run(c, world)
    got_sp, st_sp = emu_run(c, world, track_support=True)
    c.forward()
    assert rel_err(got, c.amplitudes()) < TOL and rel_err(got_sp, c.amplitudes()) < TOL
    assert st["store_side"] >= 2 and st_sp["store_side"] >= 1      # rounds ride on stores; fewer of them while qubits are still |0>
    assert st_sp["store_side"] <= st["store_side"]


def test_swap_rounds_spread_over_several_passes(monkeypatch):
    """A round of several swaps makes its pass NVLink-bound; where the victims' last uses allow, the swaps leave one per
    pass from the last passes of the step in front of the round (planner.cpp: split_swap_rounds) -- more, smaller rounds,
    never more passes than the same schedule unsplit.  Both schedules against the oracle."""
    n, world = 16, 8
    circ = OracleCircuit(n)
    circuits.random_circuit(circ, n, 500, seed=2)
    ref = OracleCircuit(n); ref.gates = list(circ.gates); ref.forward()
    stats = {}
    for split in ("1", "0"):
        monkeypatch.setenv("DVD_SPLIT_ROUNDS", split)
        got, st = emu_run(circ, world)
        got_sp, _ = emu_run(circ, world, track_support=True)
        assert rel_err(got, ref.amplitudes()) < TOL and rel_err(got_sp, ref.amplitudes()) < TOL, split
        stats[split] = st
    assert stats["1"]["store_side"] > stats["0"]["store_side"]        # more rounds ...
    assert stats["1"]["swaps"] == stats["0"]["swaps"]                  # ... of the same swaps


def test_planner_choices_record_and_replay():
    """The planners search a small portfolio (tile candidates x relabelling per plan, tail-deferral thresholds per
    schedule); the winners are recorded per gate-list structure and replayed when the same circuit comes back with new
    angles (planner.h: PlanChoices, planner.h: ChoiceMemoTable).  A replayed run must reproduce the searched schedule op for
    op -- on one rank and on emulated ranks, dense and from a reset -- and a second parameter set must replay to the
    plan a fresh search finds for it."""
    import ctypes
    from tests.helpers import gate_array
    L = emu()
    for kind, n, arg in (("random", 16, 300), ("hea", 17, 4), ("qft", 16, 0), ("layered", 15, 3)):
        c = OracleCircuit(n)
        if kind == "random":
            circuits.random_circuit(c, n, arg, seed=3)
        elif kind == "hea":
            circuits.hea(c, n, arg)
        elif kind == "qft":
            circuits.qft_like(c, n)
        else:
            circuits.layered(c, n, arg)
        arr, ng = gate_array(c)
        for g in c.gates:                      # the same circuit with other angles (a variational step)
            if g.parameter is not None:
                g.parameter = 0.37 + 1.7 * g.parameter
        arr2, ng2 = gate_array(c)
        assert ng2 == ng
        for world in (1, 2, 8):
            if n - (world.bit_length() - 1) < 12:
                continue
            for from_reset in (0, 1):
                ms = (ctypes.c_double * 2)()
                assert L.emu_plan_replay_check(n, world, arr, None, ng, from_reset, ms) == 0, (kind, n, world, from_reset, L.emu_error())
                assert L.emu_plan_replay_check(n, world, arr, arr2, ng, from_reset, ms) == 0, (kind, n, world, from_reset, L.emu_error())


def test_distributed_planner_with_fewer_local_than_rank_index_qubits():
    """More rank-index qubits wanted in one round than there are local positions to evict (n_local < log2(world)):
    the planner must bring them in over several rounds instead of evicting position -1."""
    import ctypes
    from damavand_b200 import _lib
    from tests.helpers import gate_array
    L = _lib.load()
    for n, n_local in ((3, 1), (5, 2), (4, 1), (6, 3)):
        c = OracleCircuit(n)
        for q in range(n_local, n):
            c.add_hadamard_gate(q)
        for q in range(n - 1):
            c.add_cnot_gate(q, q + 1)
        arr, ng = gate_array(c)
        perm = (ctypes.c_int32 * n)(*range(n))
        out = (ctypes.c_int32 * 4096)()
        k = L.dvd_plan_distributed_debug(n, n_local, arr, ng, perm, 1, out, 4096)
        assert k > 0, L.dvd_last_error()
        assert list(perm) == list(range(n))
        pos, n_steps = 1, out[0]
        for _ in range(n_steps):
            kind, a, b, cnt = out[pos], out[pos + 1], out[pos + 2], out[pos + 3]
            pos += 4
            if kind == 1:
                assert n_local <= a < n and 0 <= b < n_local, (n, n_local, a, b)
            for _g in range(cnt):
                assert 0 <= out[pos + 1] < n and -1 <= out[pos + 2] < n
                pos += 3
        # and the replay of that plan on 2^(n - n_local) ranks reproduces the oracle (one-gate kernel path)
        _check(c, 1 << (n - n_local))


def test_distributed_sixteen_ranks_split_their_remaps():
    """Four rank-index qubits: a swap round can involve more positions than one fused remap takes (three rank-index,
    three local), so the engine must split it -- and a restore that cycles rank-index positions composes into one
    load with bit moves between local positions."""
    n = 16
    circ = OracleCircuit(n)
    circuits.random_circuit(circ, n, 240, seed=160)
    stats = _check(circ, 16)
    assert stats["swaps"] >= 4
    circ = OracleCircuit(n)
    for q in range(n):
        circ.add_hadamard_gate(q)
    for q in (15, 14, 13, 12, 15, 13):      # every rank-index qubit in turn, twice
        circ.add_rotation_x_gate(q, 0.3 + 0.1 * q); circ.add_cnot_gate(q, (q + 5) % n)
    _check(circ, 16)


def test_distributed_small_chunks_use_simple_kernel():
    circ = OracleCircuit(8)
    circuits.random_circuit(circ, 8, 120, seed=3)
    _check(circ, 4)      # n_local = 6 < TILE_BITS
    circ = OracleCircuit(13)
    circuits.random_circuit(circ, 13, 120, seed=4)
    _check(circ, 8)      # n_local = 10


def test_forward_twice_accumulates():
    c = OracleCircuit(12); circuits.random_circuit(c, 12, 50, 9)
    st, _ = emu_run(c)
    flat = np.ascontiguousarray(st).view(np.float64).copy()
    st2, _ = emu_run(c, state=flat)
    c.forward(); c.forward()
    assert rel_err(st2, c.amplitudes()) < TOL


def test_phase_folding_run_merging_and_macro_ops():
    # RY+RZ pairs fold into one K_REALPH op per qubit and layer; H RX RY RZ runs merge into one 2x2
    c = OracleCircuit(14); g = circuits.hea(c, 14, 6); s = _check(c)
    assert s["ops"] - s["switches"] <= 14 * 6 + 4 * s["passes"]          # one op per (qubit, layer) + leftovers
    c = OracleCircuit(13); g = circuits.layered(c, 13, 4); s = _check(c)
    assert s["ops"] - s["switches"] <= 13 * 4 + 4 * s["passes"]          # H RX RY RZ -> one general 2x2
    c = OracleCircuit(16); g = circuits.qft_like(c, 16); s = _check(c)
    assert s["ops"] - s["switches"] <= 16 + 8                              # one twiddle+Hadamard op per qubit


def _cross_pass_phase_circuit(n):
    rng = np.random.default_rng(5)
    c = OracleCircuit(n)
    for q in range(n):
        c.add_hadamard_gate(q)
    for a in range(n):
        for b in range(a + 1, n, 3):
            phi = float(rng.random())
            c.add_rotation_z_gate(b, phi); c.add_cnot_gate(a, b); c.add_rotation_z_gate(b, -phi); c.add_cnot_gate(a, b)
    for q in range(n - 1, -1, -1):
        c.add_rotation_x_gate(q, 0.3 + 0.05 * q)
        c.add_pauli_x_gate((q + 5) % n, False)
    for q in range(n):
        c.add_rotation_z_gate(q, 0.1 * q); c.add_hadamard_gate(q)
    return c


def test_phases_deferred_across_passes():
    # controlled phases between qubits that sit in different passes' tiles, with non-diagonal gates on both
    # sides: the term is carried over and must be emitted before the later Hadamard / RX hits its qubit
    s = _check(_cross_pass_phase_circuit(16))
    assert s["passes"] >= 2
    _check(_cross_pass_phase_circuit(16), 4)
    _check(_cross_pass_phase_circuit(14), 2)


def test_distributed_list_scheduling_needs_few_swaps():
    # a layered ansatz on 8 ranks: the whole light cone of a layout runs before any swap is paid for
    n = 16
    c = OracleCircuit(n); circuits.hea(c, n, 5)
    s = _check(c, 8)
    assert s["swaps"] <= 9          # 3 rank-index qubits in, 3 back (+ slack); one swap per gate would be 15+


@pytest.mark.parametrize("world", [1, 2, 4])
def test_support_tracking_partial_circuits(world):
    """Circuits that leave most qubits in |0>: the engine stores only what the touched qubits span, launches only
    the tiles that can be populated and fills in the implied zeros at the first observation."""
    n = 16
    rng = np.random.default_rng(world)
    cases = []
    c = OracleCircuit(n); cases.append(c)                                   # nothing but diagonal gates
    c.add_rotation_z_gate(3, 0.4); c.add_pauli_z_gate(15, False)
    c = OracleCircuit(n); cases.append(c)                                   # X moves the single populated state around
    c.add_pauli_x_gate(15, False); c.add_pauli_x_gate(0, False); c.add_cnot_gate(15, 7); c.add_cnot_gate(3, 8)
    c = OracleCircuit(n); cases.append(c)                                   # one superposed high qubit, controls on it
    c.add_hadamard_gate(14); c.add_cnot_gate(14, 2); c.add_cnot_gate(9, 5); c.add_rotation_y_gate(13, 0.3)
    c = OracleCircuit(n); cases.append(c)                                   # support grows a few qubits at a time
    for q in rng.permutation(n)[:9]:
        c.add_rotation_x_gate(int(q), 0.1 + 0.05 * q); c.add_rotation_z_gate(int(q), 0.2)
        c.add_cnot_gate(int(q), int((q + 5) % n))
    c = OracleCircuit(n); cases.append(c)                                   # GHZ chain from the top qubit down
    c.add_hadamard_gate(n - 1)
    for q in range(n - 1, 0, -1):
        c.add_cnot_gate(q, q - 1)
    for c in cases:
        _check(c, world)


@pytest.mark.parametrize("kind", ["layered", "hea", "qft", "random"])
def test_replay_against_committed_fixtures(kind):
    """Planner + kernel logic (CPU replay, dense and support-tracking forms) against tests/golden/oracle_regression.npz."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join("tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    fx = np.load("tests/golden/oracle_regression.npz")
    c = mg.regression_case(kind, OracleCircuit)
    for track in (False, True):
        got, _ = emu_run(c, 1, track_support=track)
        assert rel_err(got, fx[f"{kind}_amplitudes"]) < TOL


@pytest.mark.parametrize("world", [1, 2])
def test_tile_relabelling_option(monkeypatch, world):
    """PlanOptions::relabel (experimental, off by default): the qubit held in the pinned low tile positions trades
    places with hotter tile qubits at the end of a pass.  Fewer passes on chain-like circuits, same amplitudes, and
    the layout is back to identity at the end (the replay compares in the reference's index order)."""
    n = 18
    def build():
        c = OracleCircuit(n)
        circuits.hea(c, n, 6)
        c.add_hadamard_gate(0); c.add_cnot_gate(0, n - 1); c.add_rotation_x_gate(1, 0.3)
        return c
    monkeypatch.setenv("DVD_RELABEL", "0")
    base, st0 = emu_run(build(), world)
    monkeypatch.setenv("DVD_RELABEL", "1")
    got, st1 = emu_run(build(), world)
    got_sp, _ = emu_run(build(), world, track_support=True)
    ref = build(); ref.forward()
    assert rel_err(base, ref.amplitudes()) < TOL
    assert rel_err(got, ref.amplitudes()) < TOL
    assert rel_err(got_sp, ref.amplitudes()) < TOL
    assert st1["passes"] <= st0["passes"], (st0, st1)
    if world == 1:
        assert st1["passes"] < st0["passes"], (st0, st1)
