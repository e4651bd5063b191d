#!/usr/bin/env python
"""bench.py -- headline benchmark of the statevector hot path (BASELINE.json `metric`).

    python bench.py --gpus 1 --steps K --warmup W          # 30-qubit QFT-style circuit, 1 GPU (cfg 3)
    torchrun ... bench.py --gpus N --steps K --warmup W    # 32-qubit random circuit, strong scaling (cfg 4)
    python bench.py --impl reference ...                   # the CPU `multithreading` path (oracle port)

A step = one pass of the hot path over one circuit: every gate applied (forward) to a DENSE state.
`value` = gates/s with everything resident in HBM, timed with CUDA events on the engine's stream; nothing is known
to be zero, every pass reads and writes every amplitude.  `from_reset` = the same circuit as the reference's API
runs it (reset to |0..0> + forward), where the engine's support tracking skips what is zero by construction.
`e2e`  = the same metric through the public Python API with host buffers: gate list in (H2D), reset + forward,
1000-shot sample and expectation values out (D2H).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

WORKLOAD_DESC = {
    "qft30": "30-qubit QFT-style circuit (H + RZ/CNOT/RZ/CNOT/RZ controlled-phase decomposition, target-major), 2205 gates, 16 GiB fp64 state",
    "hea28": "28-qubit hardware-efficient ansatz, 50 layers (RY,RZ + CNOT chain), 4150 gates, 4 GiB",
    "random32": "32-qubit random circuit over {H,RX,RY,RZ,CNOT}, 640 gates, 64 GiB total (strong scaling)",
    "hea34": "34-qubit hardware-efficient ansatz, 10 layers, 1010 gates, 256 GiB total",
    "layered20": "20-qubit layered circuit (H + RX/RY/RZ + CNOT ring, 10 layers), 1000 gates",
}
WORKLOAD_QUBITS = {"qft30": 30, "hea28": 28, "random32": 32, "hea34": 34, "layered20": 20}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs about a second before its first sample, so it is started before the warm-up; stop(t0, t1) keeps
    the samples whose timestamp lies inside the timed region [t0, t1] (datetime, host clock).  When the region is shorter
    than the 100 ms sampling period, the samples taken under load during the warm-up right before it stand in
    (`window` says which)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @classmethod
    def summarise(cls, lines, t0=None, t1=None):
        """lines: nvidia-smi csv rows (strings).  Returns the clocks dict of the bench contract."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": None}
        rows = []
        for line in lines:
            r = [x.strip() for x in line.strip().split(",")]
            if len(r) < 10:
                continue
            ts = None
            for fmt in ("%Y/%m/%d %H:%M:%S.%f", "%Y/%m/%d %H:%M:%S", "%Y-%m-%d %H:%M:%S.%f"):
                try:
                    ts = datetime.datetime.strptime(r[0], fmt)
                    break
                except ValueError:
                    pass
            try:
                power = float(r[4])
            except ValueError:
                power = 0.0
            try:
                # an unparsable timestamp keeps the row (as "before the region"): clocks are still worth reporting
                rows.append((ts or datetime.datetime.min, float(r[2]), float(r[3]), power, r[6:10]))
            except ValueError:
                continue
        if not rows:
            return out
        inside = [x for x in rows if t0 is not None and t1 is not None and t0 <= x[0] <= t1]
        if inside:
            use, window = inside, "timed region"
        else:
            # region shorter than the sampling period: fall back to the samples under load (power within 70 % of the
            # highest seen) taken before its end, i.e. during the warm-up of the same workload
            before = [x for x in rows if t1 is None or x[0] <= t1] or rows
            pmax = max(x[3] for x in before)
            use = [x for x in before if x[3] >= 0.7 * pmax] or before
            window = "under load before the end of the timed region (region shorter than the sampling period)"
        reasons = set()
        for x in use:
            for k, name in enumerate(cls.NAMES):
                if x[4][k].lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=float(np.median([x[1] for x in use])), sm_max_mhz=float(max(x[2] for x in use)),
                   reasons=sorted(reasons), samples=len(use), window=window)
        return out

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return self.summarise([])
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            lines = open(self.path).read().splitlines()
            os.unlink(self.path)
        except Exception:
            return self.summarise([])
        return self.summarise(lines, t0, t1)


def build_workload(circ, name):
    from damavand_b200 import circuits
    return circuits.workload(name)[1](circ)


def workload_qubits(name):
    from damavand_b200 import circuits
    return circuits.workload(name)[0]


def workload_desc(name):
    return WORKLOAD_DESC.get(name, name)


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's restatement of the reference `multithreading` method on a bounded sample
# --------------------------------------------------------------------------------------------------
CPU_MAX_QUBITS = 30   # 16 GiB state + 16 GiB clone, ~1.5 s per gate: keeps a CPU step of 2 gates at a few seconds


def cpu_sample_gates(name: str, n: int, mem_gib: float):
    """Qubit count of the CPU sample: what the host can hold (state + clone = 32 B/amp), at most CPU_MAX_QUBITS."""
    n_cpu = min(n, CPU_MAX_QUBITS)
    need = 32.0 * (1 << n_cpu) / 2**30
    while need + 4 > mem_gib and n_cpu > 20:
        n_cpu -= 1
        need /= 2
    return n_cpu


def host_mem_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return float(line.split()[1]) / 2**20
    except Exception:
        pass
    return 16.0


def run_cpu(name: str, steps: int, warmup: int, gates_per_step: int):
    """Time the CPU path: `steps` steps of `gates_per_step` consecutive gates of the workload's circuit.
    Returns (gates_per_sec, seconds_per_step, description, cores)."""
    from oracle.oracle import OracleCircuit
    from damavand_b200 import circuits
    n = workload_qubits(name)
    n_cpu = cpu_sample_gates(name, n, host_mem_gib())
    o = OracleCircuit(n_cpu)
    if n_cpu == n:
        circuits.workload(name)[1](o)
    else:   # same generator, fewer qubits (host RAM cannot hold state + clone)
        import re
        kind = re.match(r"[a-z]+", name).group(0)
        circuits.workload(f"{kind}{n_cpu}")[1](o)
    all_gates = [g for i, g in enumerate(o.gates) if i not in set(o.observables)]
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    total_needed = (steps + warmup) * gates_per_step
    # cycle through the circuit's gates in order
    times = []
    k = 0
    for s in range(steps + warmup):
        o.gates = [all_gates[(k + i) % len(all_gates)] for i in range(gates_per_step)]
        o.observables = []
        k += gates_per_step
        t0 = time.perf_counter()
        o.forward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    sec = float(np.sum(times))
    # the CPU path streams the whole state once per gate (clone + update), so its cost per gate is proportional to
    # 2^n: a sample taken at n_cpu < n qubits is scaled by 2^(n_cpu - n) to the workload's size and labelled so
    factor = 2.0 ** (n_cpu - n)
    gps = steps * gates_per_step / sec * factor
    scale = "" if n_cpu == n else (f", timed at {n_cpu} qubits (same generator) and scaled by 2^{n_cpu - n} to {n} qubits: "
                                   f"the per-gate cost of clone + update is linear in the state size")
    desc = (f"{steps} steps x {gates_per_step} consecutive gates of the {name} circuit{scale}; "
            f"oracle port of circuit_multithreading.rs:9-54 (clone + per-amplitude update), OpenMP {cores} threads")
    return gps, sec / steps / factor, desc, cores, n_cpu


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload or ("qft30" if args.gpus == 1 else "random32")
    n = workload_qubits(name)
    gates_per_step = args.cpu_gates_per_step or (2 if n >= 30 else 4 if n >= 28 else 50)
    gps, sec_step, desc, cores, n_cpu = run_cpu(name, args.steps, args.warmup, gates_per_step)
    factor = 2.0 ** (n - n_cpu)           # sec_step is already scaled to the workload's size
    measured_ms = sec_step * 1e3 / factor
    cfg = {"workload": f"{name}: {workload_desc(name)}", "cpu_qubits": n_cpu, "gates_per_step": gates_per_step}
    if n_cpu != n:
        # the host cannot hold 2 x 16 * 2^n B: the step is MEASURED at n_cpu qubits; `value` is that rate scaled to the
        # workload's size (cost per gate of clone + update is linear in the state size) -- an extrapolation, labelled so
        cfg["extrapolated"] = {"measured_at_qubits": n_cpu, "workload_qubits": n, "factor": factor,
                               "ms_per_step_at_workload_size": sec_step * 1e3,
                               "note": "ms_per_step is the measured time of one step of the bounded sample at measured_at_qubits; "
                                       "value = measured gates/s / factor"}
    line = {
        "impl": "reference", "metric": "gates_per_sec", "value": gps, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": measured_ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": gps, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": gps, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Ranks:
    """The torch.distributed plumbing the harness needs (nothing on the data path): barrier and max over ranks."""

    def __init__(self, gpus: int):
        self.gpus = gpus
        self.rank = int(os.environ.get("RANK", "0"))

    def barrier(self, circ=None):
        if self.gpus > 1:
            import torch
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()
        if circ is not None:
            circ.synchronize()

    def max(self, x: float) -> float:
        if self.gpus == 1:
            return x
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        return self.max(0.0 if ok else 1.0) == 0.0


def parity_check(ranks: Ranks, jit: int, cases=None):
    """The CUDA path against the CPU oracle (oracle/ is the CHECKER here, never the thing measured) on this run's own
    ranks, before anything is timed: this rank's chunk of the amplitudes (max-relative and l2-relative error), the
    allreduced norm and exact <Z_q>, and the sampler's indices bit for bit under injected uniforms (world > 1: against
    the restatement of sample_distributed, circuit_distributed.rs:42-129).  Returns the `parity` object of the JSON line."""
    from damavand_b200 import Circuit, circuits
    from oracle import oracle
    from oracle.oracle import OracleCircuit
    world, rank = ranks.gpus, ranks.rank
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(max(1, cores // world))     # every rank runs the oracle itself
    method = "gpu" if world == 1 else "distributed_gpu"
    cases = cases or [(22, "random", 300), (22, "qft", 0), (23, "hea", 4)]
    out = {"ok": True, "tolerance": 1e-12, "max_rel_err": 0.0, "l2_rel_err": 0.0, "ez_err": 0.0, "norm_err": 0.0,
           "samples_ok": True, "shots": 5000, "cases": [], "jit": bool(jit),
           "oracle": "oracle/ restatement of circuit_multithreading.rs:9-54 (checker only)"}
    for n, kind, arg in cases:
        g = Circuit(n, method)
        if jit:
            g.set_jit(2)       # compile on first use: the specialised kernels are what is checked
        o = OracleCircuit(n)
        for c in (g, o):
            if kind == "random":
                circuits.random_circuit(c, n, arg, seed=n)
            elif kind == "qft":
                circuits.qft_like(c, n)
            else:
                circuits.hea(c, n, arg)
            for q in range(n):
                c.add_pauli_z_gate(q, True)
        o.forward()
        full = o.amplitudes()
        chunk = (1 << n) // world
        want = full[rank * chunk:(rank + 1) * chunk]
        p = o.measure_np()
        idx = np.arange(p.size)
        ez_want = np.array([(p * (1 - 2.0 * ((idx >> q) & 1))).sum() for q in range(n)])
        shots = out["shots"]
        u = np.random.default_rng(1235).random(2 * shots)
        if world > 1:
            s_want = oracle.sample_distributed(p, world, u[:shots], u[shots:], "tree")
        else:
            s_want = np.asarray(o.sample(shots, uniforms=u[:shots], mode="tree"), dtype=np.uint64)
        errs = []
        for rep in range(2):          # from a reset (support-tracked passes), then reset again (plan cache, warm kernels)
            g.reset_amplitudes()
            g.forward()
            got = g.state_numpy()
            errs.append((float(np.abs(got - want).max() / np.abs(full).max()),
                         float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300))))
        err, l2 = max(e[0] for e in errs), max(e[1] for e in errs)
        ez_err = float(np.abs(g.expectation_z() - ez_want).max())
        norm_err = abs(g.norm() - 1.0)
        s = g.sample_numpy(shots, u if world > 1 else u[:shots])
        s_ok = bool((s == s_want).all())
        st = g.stats()
        ok = err < 1e-12 and l2 < 1e-12 and ez_err < 1e-12 and norm_err < 1e-12 and s_ok
        ok = ranks.all_ok(ok)
        err, l2, ez_err, norm_err = ranks.max(err), ranks.max(l2), ranks.max(ez_err), ranks.max(norm_err)
        s_ok = ranks.all_ok(s_ok)
        out["cases"].append({"n": n, "circuit": kind, "max_rel_err": err, "l2_rel_err": l2, "ez_err": ez_err, "norm_err": norm_err,
                             "samples_ok": s_ok, "global_swaps": int(st["global_swaps"]), "fused_remap_passes": int(st.get("remap_passes", 0)),
                             "store_side_remap_passes": int(st.get("store_remap_passes", 0)),
                             "passes": int(st["tile_passes"]), "jit_launches": int(st["jit_launches"]), "ok": ok})
        out["ok"] = out["ok"] and ok
        out["samples_ok"] = out["samples_ok"] and s_ok
        for k, v in (("max_rel_err", err), ("l2_rel_err", l2), ("ez_err", ez_err), ("norm_err", norm_err)):
            out[k] = max(out[k], v)
        g.close()
    out["n"] = sorted({c["n"] for c in out["cases"]})
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return out


def measure(circ, n_gates, args, ranks: Ranks, jit: int, device: int, sample_clocks: bool):
    """Warm-up (incl. run-time compilation and kernel-form selection), then the two timed regions:
    `from_reset` = --steps x (reset + forward), `dense` = --steps x forward on the dense state left behind."""
    import datetime

    def one_step():
        circ.reset_amplitudes()
        circ.forward_async()

    if jit:
        circ.set_jit(1)
    sampler = ClockSampler(device)     # started before the warm-up: nvidia-smi needs ~1 s before its first sample
    if sample_clocks and ranks.rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    if jit:
        # warm-up includes the run-time compilation of the pass kernels of this circuit's structure (sparse step and
        # dense forward): wait for the background compiler, then run each once more so that the modules are loaded
        circ.synchronize(); circ.jit_wait()
        one_step(); circ.forward_async(); circ.synchronize(); circ.jit_wait()
        # ... and the kernel form of every pass structure is chosen by timing its candidates on dense launches
        # (csrc/jit_rt.cpp): keep warming up until nothing is being measured any more
        last, idle = None, 0
        for _ in range(12):
            one_step(); circ.forward_async(); circ.synchronize()
            left = circ.jit_info()["tuning"]
            idle = idle + 1 if left == last else 0
            last = left
            if left == 0 or idle >= 2:       # (structures of another circuit of this process may stay half-measured)
                break
    ranks.barrier(circ)
    circ.stats_reset()
    ranks.barrier(circ)
    t_wall0 = time.perf_counter()
    circ.timer_begin()
    for _ in range(args.steps):
        one_step()
    ms = circ.timer_end()
    ranks.barrier(circ)
    t_wall = time.perf_counter() - t_wall0
    st = circ.stats()
    jit_cfg = dict(circ.jit_info(), launches_in_from_reset_region=int(st.get("jit_launches", 0))) if jit else False
    ms = ranks.max(ms)

    # The same circuit applied to a DENSE state (the one the timed steps left behind, no reset): every pass reads and
    # writes every amplitude, nothing is known to be zero.  This is the headline region and the timing the roofline of
    # the dominant kernel is computed from.
    circ.synchronize()
    circ.forward_async(); circ.synchronize()          # the state of the last step is only partly dense for some circuits
    circ.stats_reset()
    ranks.barrier(circ)
    reps = max(1, args.steps)                          # EXACTLY --steps forwards
    t_region0 = datetime.datetime.now()
    t_wall0 = time.perf_counter()
    circ.timer_begin()
    for _ in range(reps):
        circ.forward_async()
    fwd_ms = circ.timer_end() / reps
    ranks.barrier(circ)
    t_wall_dense = time.perf_counter() - t_wall0
    clocks = sampler.stop(t_region0, datetime.datetime.now()) if (sample_clocks and ranks.rank == 0) else None
    fwd_ms = ranks.max(fwd_ms)
    st1 = {k: (v / reps if isinstance(v, (int, float)) else v) for k, v in circ.stats().items()}
    if jit:
        jit_cfg = dict(jit_cfg, launches_per_dense_forward=st1.get("jit_launches", 0), final=circ.jit_info())
    return {"from_reset_ms": ms / args.steps, "from_reset_stats": st, "dense_ms": fwd_ms, "dense_stats": st1, "reps": reps,
            "clocks": clocks, "jit": jit_cfg, "wall_dense": t_wall_dense, "wall_from_reset": t_wall,
            "value": n_gates / (fwd_ms / 1e3), "from_reset_value": n_gates / (ms / args.steps / 1e3)}


def roofline_of(m, name, n_local, n_gates, gpus):
    """`roofline` object of the bench contract for the dominant kernel (the fused pass), from the dense region."""
    st1, fwd_ms = m["dense_stats"], m["dense_ms"]
    launches = st1["tile_passes"] + st1["simple_passes"]
    if launches <= 0:
        return None
    peak, peak_src, _ = measured_peaks()
    fwd_s = fwd_ms / 1e3
    hbm_gbs = st1["pass_bytes"] / fwd_s / 1e9
    alg_gbs = st1["gate_algorithmic_bytes"] / fwd_s / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{name}_g{gpus}", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    jitted = st1.get("jit_launches", 0) >= st1["tile_passes"] > 0
    kernel = ("dvd_pass_static (k_tile_pass specialised at run time for the pass's op structure, csrc/jit.cpp)" if jitted
              else "k_tile_pass (interpreter form)" if st1["tile_passes"] else "k_simple_gate")
    r = {"bound": "hbm", "kernel": kernel,
         # bytes one launch of the pass kernel really moves through HBM: it reads and writes every local amplitude
         # once, 32 * 2^n_local B, however many gates it applies (ncu: dram read + write per launch = `traffic`)
         "achieved": hbm_gbs, "peak": peak, "unit": "GB/s", "frac": hbm_gbs / peak, "traffic": traffic, "peak_source": peak_src,
         "launches_per_circuit": launches, "avg_launch_ms": fwd_ms / launches,
         "hbm_bytes_per_launch": st1["pass_bytes"] / launches,
         "hbm_pass_gbs": hbm_gbs, "hbm_pass_frac": hbm_gbs / peak,
         # SURVEY 8(d) counts one HBM pass per gate (32 * 2^n B, 16 * 2^n if controlled), as the reference executes them;
         # a fused launch applies gates_per_launch of them in one pass, so this figure exceeds the peak
         "achieved_algorithmic": alg_gbs, "algorithmic_frac": alg_gbs / peak,
         "algorithmic_bytes_per_launch": st1["gate_algorithmic_bytes"] / launches, "gates_per_launch": n_gates / launches,
         "note": "frac = HBM bytes the pass kernel moves per launch / mean launch duration (CUDA events over the timed region, "
                 "includes NVLink exchanges at N > 1) / measured copy peak; achieved_algorithmic is the per-gate accounting of "
                 "SURVEY 8(d) and is inflated by fusion"}
    # Second roofline: the fp64 pipe.  The planner counts the DADD / DMUL / DFMA each thread executes per pass (an estimate
    # from the op list: general 2x2 = 16 per pair, real = 8, Hadamard = 4, complex multiply = 4); the peak is the DFMA
    # issue rate measured on B200 (profiles/r1_microbench2.log: 33.5 TFLOP/s = 16.75e12 fp64 instructions per second).
    # Gate-dense passes (ansatz layers, random circuits) are bound by this pipe, not by HBM.
    if st1.get("pass_fp64_instr"):
        rate = st1["pass_fp64_instr"] / fwd_s
        r["fp64"] = {"instr_per_s": rate, "peak_instr_per_s": 16.75e12, "frac": rate / 16.75e12,
                     "instr_per_amplitude_per_pass": st1["pass_fp64_instr"] / launches / (st1["pass_bytes"] / launches / 32.0),
                     "bound_ms_per_circuit": st1["pass_fp64_instr"] / 16.75e12 * 1e3,
                     "hbm_bound_ms_per_circuit": st1["pass_bytes"] / (peak * 1e9) * 1e3,
                     "note": "planner estimate of executed fp64 instructions / time / measured DFMA issue peak; the circuit cannot run "
                             "faster than max(bound_ms_per_circuit, hbm_bound_ms_per_circuit)"}
    if st1["global_swaps"]:
        # NVLink side: fused-remap passes pull (1 - 2^-k) of a chunk from partner ranks while they run (the same
        # amount leaves through the other direction of the links); stand-alone exchanges are timed on their own
        nv_ms = st1.get("remap_ms", 0.0) + st1.get("swap_ms", 0.0)
        nv_bytes = float(st1["swap_bytes_sent"])
        plain = launches - st1.get("remap_passes", 0)
        plain_ms = (fwd_ms - nv_ms) / plain if plain > 0 else None
        r["nvlink"] = {"bytes_per_dir": nv_bytes, "ms": nv_ms, "gbs_per_dir": nv_bytes / (nv_ms / 1e3) / 1e9 if nv_ms > 0 else None,
                       "frac_of_900": nv_bytes / (nv_ms / 1e3) / 1e9 / 900.0 if nv_ms > 0 else None,
                       "frac_of_measured_770": nv_bytes / (nv_ms / 1e3) / 1e9 / 770.0 if nv_ms > 0 else None,
                       "global_swaps": st1["global_swaps"], "fused_remap_passes": st1.get("remap_passes", 0),
                       "store_side_remap_passes": st1.get("store_remap_passes", 0),
                       "avg_store_side_pass_ms": st1.get("store_remap_ms", 0.0) / st1["store_remap_passes"] if st1.get("store_remap_passes") else None,
                       "avg_load_side_pass_ms": (st1.get("remap_ms", 0.0) - st1.get("store_remap_ms", 0.0)) / (st1["remap_passes"] - st1.get("store_remap_passes", 0))
                                                if st1.get("remap_passes", 0) > st1.get("store_remap_passes", 0) else None,
                       "avg_fused_pass_ms": st1.get("remap_ms", 0.0) / st1["remap_passes"] if st1.get("remap_passes") else None,
                       "avg_plain_pass_ms": plain_ms,
                       "what": "per rank and per forward; ms = device time (CUDA events) of the passes whose load -- or, for the layout "
                               "restore that ends the gate list, whose store -- carries the swaps (they also apply their gates: the "
                               "exchange is their read / write) plus any stand-alone exchange"}
    return r


def cfg5_point(args, ranks: Ranks, jit: int):
    """BASELINE config 5 on 8 ranks: 34-qubit ansatz (10 layers), forward from a reset, 10^5-shot distributed sample,
    expectation values; parity flag from a 22-qubit twin of the same circuit against the oracle."""
    from damavand_b200 import Circuit, circuits
    out = {"workload": workload_desc("hea34"), "n_gpus": ranks.gpus}
    twin = parity_check(ranks, jit, cases=[(22, "hea", 10)])
    out["parity_twin"] = {k: twin[k] for k in ("ok", "max_rel_err", "l2_rel_err", "samples_ok", "n")}
    n = 34
    c = Circuit(n, "distributed_gpu")
    ng = circuits.hea(c, n, 10, observables=True)
    if jit:
        c.set_jit(1)
    for _ in range(2):
        c.reset_amplitudes(); c.forward_async()
    c.synchronize()
    if jit:
        c.jit_wait()
        last, idle = None, 0
        for _ in range(8):       # kernel forms settle on dense launches (most passes of the ansatz are dense after two layers)
            c.reset_amplitudes(); c.forward_async(); c.synchronize()
            left = c.jit_info()["tuning"]
            idle = idle + 1 if left == last else 0
            last = left
            if left == 0 or idle >= 2:
                break
    ranks.barrier(c)
    c.stats_reset()
    k = max(1, min(args.steps, 3))
    c.timer_begin()
    for _ in range(k):
        c.reset_amplitudes(); c.forward_async()
    fwd = ranks.max(c.timer_end() / k)
    st = {a: (b / k if isinstance(b, (int, float)) else b) for a, b in c.stats().items()}
    shots = 100000
    u = np.random.default_rng(1235).random(2 * shots)
    ranks.barrier(c)
    t0 = time.perf_counter()
    s = c.sample_numpy(shots, u)
    t_sample = ranks.max(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ev = c.extract_expectation_values_numpy(s)
    ev_mean = ev.mean(axis=0)
    t_ev = ranks.max(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ez = c.expectation_z()
    t_ez = ranks.max(time.perf_counter() - t0)
    norm = c.norm()
    out.update({"gates": ng, "forward_ms": fwd, "gates_per_sec": ng / (fwd / 1e3), "passes": st["tile_passes"], "global_swaps": st["global_swaps"],
                "fused_remap_passes": st.get("remap_passes", 0), "store_side_remap_passes": st.get("store_remap_passes", 0),
                "nvlink_bytes_per_dir": st["swap_bytes_sent"],
                "sample_shots": shots, "sample_ms": t_sample * 1e3, "extract_expectation_values_ms": t_ev * 1e3,
                "expectation_z_ms": t_ez * 1e3, "norm_err": abs(norm - 1.0),
                "sampled_vs_exact_z_max_abs_diff": float(np.abs(ev_mean - ez).max()),
                "what": "reset + forward timed with CUDA events (max over ranks); sample = tree build + rank draw + per-rank descent + "
                        "allreduce, wall clock with host buffers; sampled <Z> against the exact reduction as a statistical cross-check "
                        "(expected |diff| ~ 1/sqrt(shots) = 0.003)"})
    c.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--cpu-gates-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scaling-point", action="store_true")
    ap.add_argument("--no-single-gate", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison that precedes the timed regions")
    ap.add_argument("--no-cfg5", action="store_true", help="N = 8 only: skip the 34-qubit config-5 point")
    ap.add_argument("--unfused", action="store_true", help="one kernel per gate (no fusion) for comparison")
    ap.add_argument("--jit", type=int, default=None, help="structure-specialised kernels: 0 off, 1 on (compiled during warm-up)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    from damavand_b200 import Circuit, distributed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks"}))
            return 2
    ranks = Ranks(args.gpus)
    rank = ranks.rank
    name = args.workload or ("qft30" if args.gpus == 1 else "random32")
    n = workload_qubits(name)
    method = "gpu" if args.gpus == 1 else "distributed_gpu"
    if args.gpus > 1:
        distributed.initialize("nccl")
    device = int(os.environ.get("LOCAL_RANK", "0"))
    # run-time specialised pass kernels (validated against the oracle in the parity leg below and in tests/)
    jit = args.jit if args.jit is not None else int(os.environ.get("DVD_BENCH_JIT", "1"))

    # ---- parity first: a fast step that computes the wrong state is not a result -------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_check(ranks, jit)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"error": "parity check against the oracle failed", "parity": parity}), flush=True)
            return 3

    circ = Circuit(n, method)
    if args.unfused:
        circ.set_unfused(True)
    n_gates = build_workload(circ, name)
    m = measure(circ, n_gates, args, ranks, jit, device, sample_clocks=True)
    st, st1 = m["from_reset_stats"], m["dense_stats"]
    # the state the timed forwards left behind is still a unit vector (one reduction pass, after the timed region)
    norm = circ.norm()
    sanity = {"norm_after_timed_region": norm, "ok": abs(norm - 1.0) < 1e-10}
    n_local = n - int(np.log2(args.gpus))
    roofline = roofline_of(m, name, n_local, n_gates, args.gpus)
    peak = measured_peaks()[0]

    # the pass kernel with ONE gate per pass: gate application as the reference does it (one pass over HBM per
    # gate), on the state the circuit just produced.  This is the number that compares with "gate application at
    # >= 70 % of the HBM peak": algorithmic bytes of the gate / time of its pass.
    if roofline is not None and args.gpus == 1 and not args.no_single_gate:
        saved_gates, saved_obs = circ.gates, circ.observables
        single = {}
        for label, build_one in (("H(q=n-1): uncontrolled, highest stride", lambda c: c.add_hadamard_gate(n - 1)),
                                 ("RX(q=0): uncontrolled, lowest stride", lambda c: c.add_rotation_x_gate(0, 0.3)),
                                 ("CNOT(c=n-1,t=n-2): controlled, half the amplitudes", lambda c: c.add_cnot_gate(n - 1, n - 2))):
            circ.gates, circ.observables = [], []
            build_one(circ)
            reps = 6
            for _ in range(4 if jit else 1):                  # warm-up (non-zero input: no lazy reset involved; kernel form settles)
                circ.forward_async(); circ.synchronize()
                if jit:
                    circ.jit_wait()
            circ.timer_begin()
            for _ in range(reps):
                circ.forward_async()
            t_ms = circ.timer_end() / reps
            alg = (16.0 if "CNOT" in label else 32.0) * (1 << n)
            single[label] = {"ms": t_ms, "algorithmic_gbs": alg / (t_ms / 1e3) / 1e9,
                             "hbm_gbs": 32.0 * (1 << n) / (t_ms / 1e3) / 1e9,
                             "hbm_frac": 32.0 * (1 << n) / (t_ms / 1e3) / 1e9 / peak}
        circ.gates, circ.observables = saved_gates, saved_obs
        roofline["single_gate_pass"] = single
        roofline["single_gate_note"] = ("one gate per launch of the same kernel (what the reference's per-gate kernel does): "
                                        "hbm_gbs = 32*2^n B / pass time; a CNOT pass still reads and writes every tile")

    # e2e through the public API with host buffers: gates in, samples + expectation values out
    e2e = None
    if not args.no_e2e:
        shots = 1000
        if not circ.observables:
            for q in range(n):
                circ.add_pauli_z_gate(q, True)
        per_shot = 2 if args.gpus > 1 else 1
        u = np.random.default_rng(1235).random(shots * per_shot)
        k_e2e = max(1, min(args.steps, 3))
        ranks.barrier(circ)
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            circ.reset_amplitudes()
            circ.forward()
            s = circ.sample_numpy(shots, u)
            ev = circ.extract_expectation_values_numpy(s)
            float(ev.mean())
        ranks.barrier(circ)
        dt = ranks.max((time.perf_counter() - t0) / k_e2e)
        h2d = 80 * n_gates + 8 * shots * per_shot + 8 * shots + 4 * n
        d2h = 8 * shots + 8 * shots * n
        e2e = {"value": n_gates / dt, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3,
               "what": "reset + Circuit.forward() (plan, upload gate list, fused passes) + sample(1000) + extract_expectation_values, host buffers, wall clock"}
    circ.close()
    circ = None

    # Strong scaling (cfg 4): the N > 1 lines run random32 and must be divided by the 1-GPU rate of the SAME workload,
    # measured the same way (same step, same kernels, --steps / --warmup).  N = 1: it travels as `scaling_point` (the
    # headline line itself is cfg 3).  N > 1: rank 0 measures it on its own GPU right here -- the other ranks wait --
    # so that every line carries its own efficiency and nobody divides qft30 by random32.
    scaling_point = strong_scaling = None
    want_base = (args.gpus == 1 and args.workload is None and not args.no_scaling_point) or \
                (args.gpus > 1 and not args.no_scaling_point)
    if want_base:
        sp_name = "random32" if args.gpus == 1 else name
        base = None
        if rank == 0:
            try:
                c2 = Circuit(workload_qubits(sp_name), "gpu")
                ng2 = build_workload(c2, sp_name)
                m2 = measure(c2, ng2, args, Ranks(1), jit, device, sample_clocks=True)
                base = {"workload": f"{sp_name}: {workload_desc(sp_name)}", "n_gpus": 1, "value": m2["value"], "unit": "gates/s",
                        "ms_per_step": m2["dense_ms"], "from_reset_ms": m2["from_reset_ms"], "steps": args.steps, "warmup": args.warmup,
                        "passes": int(round(m2["dense_stats"]["tile_passes"])), "jit": bool(jit), "clocks": m2["clocks"],
                        "hbm_pass_frac": (roofline_of(m2, sp_name, workload_qubits(sp_name), ng2, 1) or {}).get("hbm_pass_frac"),
                        "note": "same step definition as `value` (forward on a dense state), same kernels, same --steps/--warmup"}
                c2.close()
            except Exception as e:
                base = {"workload": sp_name, "value": None, "note": f"failed: {e}"}
        ranks.barrier()
        if args.gpus == 1:
            scaling_point = base
        elif rank == 0:
            strong_scaling = {"base": base, "base_value": base.get("value"), "value": m["value"],
                              "efficiency": (m["value"] / (args.gpus * base["value"])) if base.get("value") else None,
                              "what": "value / (n_gpus x base_value): 1-GPU rate of the same workload measured by rank 0 in this run"}

    cfg5 = None
    if args.gpus == 8 and args.workload is None and not args.no_cfg5:
        try:
            cfg5 = cfg5_point(args, ranks, jit)
        except Exception as e:
            cfg5 = {"error": f"{type(e).__name__}: {e}"}

    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        try:
            gps_step = args.cpu_gates_per_step or (2 if n >= 30 else 4 if n >= 28 else 50)
            gps, sec_step, desc, cores, n_cpu = run_cpu(name, 3, 1, gps_step)
            cpu_baseline = {"value": gps, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as e:  # the CPU leg must never break the GPU number
            cpu_baseline = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        dense_state = {"value": m["value"], "unit": "gates/s", "ms_per_step": m["dense_ms"],
                       "what": "forward of the same circuit on a dense state (no reset, nothing known to be zero): every pass "
                               "moves 32*2^n_local B"}
        line = {
            "metric": "gates_per_sec", "value": m["value"], "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["dense_ms"], "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {workload_desc(name)}", "apply_method": method,
                       "step": "forward of the whole circuit on a DENSE state (no reset between steps, nothing known to be zero: "
                               "every pass reads and writes every amplitude).  `from_reset` is the same circuit as the "
                               "reference's API runs it, reset to |0..0> + forward, where the engine tracks which qubits have "
                               "left |0> and neither reads nor launches tiles that are zero by construction",
                       "l2": "state (>= 4 GiB per GPU) is far larger than the 126 MB L2; no flush needed",
                       "fused": not args.unfused, "wall_s_timed_region": m["wall_dense"], "wall_s_from_reset_region": m["wall_from_reset"],
                       "jit": m["jit"]},
            "gpu_launches": int(round(st1["kernel_launches"] * m["reps"])),
            "passes_per_circuit": int(round(st1["tile_passes"])),
            "global_swaps_per_circuit": int(round(st1["global_swaps"])),
            "from_reset": {"value": m["from_reset_value"], "unit": "gates/s", "ms_per_step": m["from_reset_ms"], "steps": args.steps,
                           "gpu_launches": int(st["kernel_launches"]),
                           "what": "reset to |0..0> + forward, timed like `value` (the step every BASELINE config describes)"},
            "dense_state": dense_state, "parity": parity, "sanity": sanity,
            "clocks": m["clocks"], "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline,
        }
        if scaling_point is not None:
            line["scaling_point"] = scaling_point
        if strong_scaling is not None:
            line["strong_scaling"] = strong_scaling
        if cfg5 is not None:
            line["cfg5_point"] = cfg5
        print(json.dumps(line), flush=True)
    if args.gpus > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0 if sanity["ok"] else 4


if __name__ == "__main__":
    sys.exit(main())
