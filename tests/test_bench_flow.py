"""bench.py's GPU arm, end to end on the CPU with a stub engine: the control flow, the JSON contract and the key
arithmetic are checked where no GPU exists (the numbers themselves are meaningless here).  A real run is what the
`-m gpu` box and the driver do."""
import io
import json
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest


class StubCircuit:
    """Mimics the parts of damavand_b200.Circuit that bench.py uses; every forward 'takes' 2 ms."""
    instances = 0

    def __init__(self, num_qubits, apply_method=None):
        self.num_qubits, self.apply_method = num_qubits, apply_method
        self.gates, self.observables = [], []
        self._st = dict(gates_applied=0, kernel_launches=0, tile_passes=0, simple_passes=0, stage_switches=0, global_swaps=0,
                        swap_bytes_sent=0, pass_bytes=0.0, gate_algorithmic_bytes=0.0, plan_cache_hits=0, jit_launches=0,
                        remap_passes=0, remap_bytes_in=0.0, remap_ms=0.0, swap_ms=0.0, pass_fp64_instr=0.0, store_remap_passes=0, store_remap_ms=0.0)
        self._t = 0.0
        self._t0 = 0.0
        self.closed = False
        self.jit = 0
        StubCircuit.instances += 1

    def _add(self, *a):
        self.gates.append(list(a))

    def add_hadamard_gate(self, q): self._add("Hadamard", q, None, None)
    def add_rotation_x_gate(self, q, t): self._add("RotationX", q, None, t)
    def add_rotation_y_gate(self, q, t): self._add("RotationY", q, None, t)
    def add_rotation_z_gate(self, q, t): self._add("RotationZ", q, None, t)
    def add_cnot_gate(self, c, t): self._add("CNOT", t, c, None)

    def add_pauli_z_gate(self, q, obs):
        self._add("PauliZ", q, None, None)
        if obs:
            self.observables.append(len(self.gates) - 1)

    add_pauli_x_gate = add_pauli_y_gate = add_pauli_z_gate

    def _alive(self):
        assert not self.closed, "bench.py used a circuit after closing it"

    def reset_amplitudes(self): self._alive()
    def set_unfused(self, f): self._alive()
    def set_jit(self, m): self._alive(); self.jit = m
    def jit_wait(self): self._alive()
    def jit_info(self): self._alive(); return dict(compiled=3, failed=0, pending=0, compile_seconds=1.0, message="", tuning=0, chosen={}, launches={})
    def synchronize(self): self._alive()
    def stats_reset(self): self._alive(); self._st = {k: type(v)(0) for k, v in self._st.items()}

    def forward_async(self):
        self._alive()
        n = len(self.gates) - len(self.observables)
        self._t += 2.0
        st = self._st
        st["gates_applied"] += n; st["kernel_launches"] += 3; st["tile_passes"] += 3
        st["pass_bytes"] += 3 * 32.0 * (1 << self.num_qubits); st["gate_algorithmic_bytes"] += n * 32.0 * (1 << self.num_qubits)
        if self.jit:
            st["jit_launches"] += 3
        return n

    forward = forward_async

    def timer_begin(self): self._alive(); self._t0 = self._t
    def timer_end(self): self._alive(); return self._t - self._t0
    def stats(self): self._alive(); return dict(self._st)
    def sample_numpy(self, shots, u): self._alive(); return np.zeros(shots, dtype=np.uint64)
    def extract_expectation_values_numpy(self, s): self._alive(); return np.ones((len(s), max(1, len(self.observables))))
    def norm(self): self._alive(); return 1.0
    def close(self): self.closed = True


@pytest.mark.parametrize("argv", [[], ["--workload", "hea28", "--no-cpu-baseline"], ["--jit", "0", "--no-scaling-point", "--no-e2e"]])
def test_gpu_arm_flow_with_stub_engine(monkeypatch, argv):
    import damavand_b200
    import bench
    monkeypatch.setattr(damavand_b200, "Circuit", StubCircuit)
    monkeypatch.setattr(bench, "run_cpu", lambda name, steps, warmup, gps: (0.5, 4.0, "stub sample", 8, 30))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    # (the oracle comparison that precedes the timed regions needs the real engine: covered by the -m gpu tests)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3", "--no-parity"] + argv)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.main() == 0
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "e2e", "cpu_baseline", "dense_state", "from_reset"):
        assert k in d, k
    assert d["metric"] == "gates_per_sec" and d["unit"] == "gates/s" and d["steps"] == 4 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["higher_is_better"] is True
    n_gates = 4150 if "hea28" in argv else 2205
    # every stub forward takes 2 ms: both step definitions give n_gates / 2 ms
    assert d["value"] == pytest.approx(n_gates / 2e-3) and d["ms_per_step"] == pytest.approx(2.0)
    assert d["from_reset"]["value"] == pytest.approx(n_gates / 2e-3)
    assert d["gpu_launches"] == 3 * 4
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    # frac is the HBM fraction: bytes the passes move (3 x 32 * 2^n per forward) / time / peak; the per-gate accounting is separate
    assert r["achieved"] == pytest.approx(3 * 32.0 * (1 << (28 if "hea28" in argv else 30)) / 2e-3 / 1e9)
    assert r["achieved_algorithmic"] == pytest.approx(n_gates * 32.0 * (1 << (28 if "hea28" in argv else 30)) / 2e-3 / 1e9)
    assert ("specialised" in r["kernel"]) == ("--jit" not in argv)
    assert d["sanity"]["ok"] is True and d["parity"] is None
    assert r["launches_per_circuit"] == 3 and r["avg_launch_ms"] == pytest.approx(2.0 / 3)
    if "--no-e2e" in argv:
        assert d["e2e"] is None
    else:
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    if "--no-cpu-baseline" in argv:
        assert d["cpu_baseline"] is None
    else:
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 8
    assert ("scaling_point" in d) == (argv == [])
    assert (d["config"]["jit"] is False) == ("--jit" in argv)


def test_clock_sampler_windowing():
    """nvidia-smi rows are filtered to the timed region by timestamp; a region shorter than the sampling period falls
    back to the samples under load before its end; throttle reasons are collected from the rows used."""
    import datetime
    import bench
    row = lambda t, sm, p, cap="Not Active": f"2026/10/17 13:56:{t:06.3f}, 0, {sm}, 1965, {p}, 0x0, Not Active, Not Active, Not Active, {cap}"
    lines = [row(1.0, 345, 180.0), row(1.1, 1965, 700.0), row(1.2, 1950, 720.0, "Active"), row(1.3, 1965, 710.0), row(1.4, 600, 200.0), "garbage"]
    T = lambda sec: datetime.datetime(2026, 10, 17, 13, 56, int(sec), int(round((sec % 1) * 1e6)))
    c = bench.ClockSampler.summarise(lines, T(1.15), T(1.35))
    assert c["samples"] == 2 and c["sm_mhz"] == 1957.5 and c["sm_max_mhz"] == 1965 and c["reasons"] == ["sw_power_cap"]
    assert c["window"] == "timed region"
    c = bench.ClockSampler.summarise(lines, T(1.31), T(1.33))          # no sample inside: samples under load before 1.33
    assert c["samples"] == 3 and c["sm_mhz"] == 1965 and "shorter" in c["window"]
    assert bench.ClockSampler.summarise([])["samples"] == 0
