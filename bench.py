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
    line = {
        "impl": "reference", "metric": "gates_per_sec", "value": gps, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_step * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {workload_desc(name)}", "cpu_qubits": n_cpu, "gates_per_step": gates_per_step},
        "cpu_baseline": {"value": gps, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": gps, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
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
    rank = int(os.environ.get("RANK", "0"))
    name = args.workload or ("qft30" if args.gpus == 1 else "random32")
    n = workload_qubits(name)
    method = "gpu" if args.gpus == 1 else "distributed_gpu"
    if args.gpus > 1:
        import torch
        import torch.distributed as dist
        distributed.initialize("nccl")

    circ = Circuit(n, method)
    if args.unfused:
        circ.set_unfused(True)
    n_gates = build_workload(circ, name)
    n_obs_gates = len(circ.observables)
    device = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier():
        if args.gpus > 1:
            import torch
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()
        circ.synchronize()

    def one_step():
        circ.reset_amplitudes()
        circ.forward_async()

    # run-time specialised pass kernels: on for the single-GPU workloads (validated on B200, profiles/r1_*_v11*);
    # the multi-GPU runs keep the interpreter kernels unless --jit 1 is given
    jit = args.jit if args.jit is not None else int(os.environ.get("DVD_BENCH_JIT", "1" if args.gpus == 1 else "0"))
    if jit:
        circ.set_jit(1)
    sampler = ClockSampler(device)     # started before the warm-up: nvidia-smi needs ~1 s before its first sample
    if rank == 0:
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
        for _ in range(12):
            one_step(); circ.forward_async(); circ.synchronize()
            if circ.jit_info()["tuning"] == 0:
                break
    barrier()
    circ.stats_reset()
    barrier()
    t_wall0 = time.perf_counter()
    circ.timer_begin()
    for _ in range(args.steps):
        one_step()
    ms = circ.timer_end()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    st = circ.stats()
    # read while the circuit is alive (the scaling-point leg below closes it)
    jit_cfg = dict(circ.jit_info(), launches_in_from_reset_region=int(st.get("jit_launches", 0))) if jit else False
    if args.gpus > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sec = ms / 1e3
    value = args.steps * n_gates / sec

    # The same circuit applied to a DENSE state (the one the timed steps left behind, no reset): every pass reads and
    # writes every amplitude, nothing is known to be zero.  This is the timing the roofline of the dominant kernel
    # (k_tile_pass) is computed from, and it is reported next to `value` as `dense_state`.
    circ.synchronize()
    circ.forward_async(); circ.synchronize()          # the state of the last step is only partly dense for some circuits
    circ.stats_reset()
    barrier()
    dense_reps = max(1, args.steps)                    # the headline region: EXACTLY --steps forwards
    import datetime
    t_region0 = datetime.datetime.now()
    t_wall0 = time.perf_counter()
    circ.timer_begin()
    for _ in range(dense_reps):
        circ.forward_async()
    fwd_ms = circ.timer_end() / dense_reps
    barrier()
    t_wall_dense = time.perf_counter() - t_wall0
    clocks = sampler.stop(t_region0, datetime.datetime.now()) if rank == 0 else None
    if args.gpus > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([fwd_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fwd_ms = float(t.item())
    st1 = {k: (v / dense_reps if isinstance(v, (int, float)) else v) for k, v in circ.stats().items()}
    dense_state = {"value": n_gates / (fwd_ms / 1e3), "unit": "gates/s", "ms_per_step": fwd_ms,
                   "what": "forward of the same circuit on a dense state (no reset, nothing known to be zero): every pass "
                           "moves 32*2^n_local B"}
    launches_fwd = st1["tile_passes"] + st1["simple_passes"]
    peak, peak_src, _ = measured_peaks()
    # algorithmic bytes per launch of the pass kernel: read + write of every local amplitude once
    n_local = n - int(np.log2(args.gpus))
    pass_bytes = 32.0 * (1 << n_local)
    roofline = None
    if launches_fwd > 0:
        fwd_s = fwd_ms / 1e3
        # a non-controlled gate = 32*2^n_local B, a controlled gate = 16*2^n_local B; a launch of the
        # pass kernel processes (gates / launches) of them.  achieved = algorithmic bytes per launch / mean launch
        # duration = total algorithmic bytes / forward time (CUDA events on the engine's stream).
        achieved = st1["gate_algorithmic_bytes"] / fwd_s / 1e9
        hbm_gbs = st1["pass_bytes"] / fwd_s / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{name}_g{args.gpus}", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_tile_pass" if st1["tile_passes"] else "k_simple_gate",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "launches_per_circuit": launches_fwd, "avg_launch_ms": fwd_ms / launches_fwd,
                    "algorithmic_bytes_per_launch": st1["gate_algorithmic_bytes"] / launches_fwd,
                    "hbm_bytes_per_launch": st1["pass_bytes"] / launches_fwd,
                    "hbm_pass_gbs": hbm_gbs, "hbm_pass_frac": hbm_gbs / peak,
                    "note": "achieved counts SURVEY 8(d) algorithmic bytes (32*2^n B per gate, 16*2^n if controlled); "
                            "it exceeds the HBM peak because one launch applies many gates while moving 32*2^n B once "
                            "(16*2^n for the first pass after a reset, which synthesises |0..0> instead of reading it). "
                            "hbm_pass_gbs = bytes the passes really move / time (what ncu's dram__bytes shows) and "
                            "hbm_pass_frac = that / peak is the fraction to compare with the 70 % target; the fused "
                            "passes are latency / fp64-pipe bound at 16 warps per SM, not HBM bound (DESIGN.md section 3)"}
        if st1["global_swaps"]:
            roofline["kernel"] += " + NVLink half-chunk swaps"
            roofline["global_swaps"] = st1["global_swaps"]
            roofline["swap_bytes_sent_per_rank"] = st1["swap_bytes_sent"]

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
            circ.forward_async(); circ.synchronize()          # warm-up (non-zero input: no lazy reset involved)
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
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            circ.reset_amplitudes()
            circ.forward()
            s = circ.sample_numpy(shots, u)
            ev = circ.extract_expectation_values_numpy(s)
            float(ev.mean())
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        if args.gpus > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = 80 * n_gates + 8 * shots * per_shot + 8 * shots + 4 * n
        d2h = 8 * shots + 8 * shots * n
        e2e = {"value": n_gates / dt, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3,
               "what": "reset + Circuit.forward() (plan, upload gate list, fused passes) + sample(1000) + extract_expectation_values, host buffers, wall clock"}

    # 1-GPU point of the strong-scaling workload the N > 1 runs use (cfg 4, random32): the headline N = 1 line is
    # the 30-qubit circuit BASELINE.json's metric names, so the scaling baseline travels as an extra key
    scaling_point = None
    if args.gpus == 1 and args.workload is None and not args.no_scaling_point:
        try:
            circ.close()
            circ = None
            sp_name = "random32"
            c2 = Circuit(workload_qubits(sp_name), "gpu")
            ng2 = build_workload(c2, sp_name)
            # same step definition and the same (interpreter) kernels as the N > 1 runs use by default
            c2.reset_amplitudes(); c2.forward_async()
            c2.forward_async(); c2.synchronize()
            c2.timer_begin()
            for _ in range(2):
                c2.forward_async()
            ms2 = c2.timer_end()
            scaling_point = {"workload": f"{sp_name}: {workload_desc(sp_name)}", "n_gpus": 1, "value": 2 * ng2 / (ms2 / 1e3),
                             "unit": "gates/s", "ms_per_step": ms2 / 2, "steps": 2, "warmup": 2, "jit": False,
                             "note": "same step definition as `value` (forward on a dense state), interpreter kernels as in "
                                     "the default N > 1 runs; divide an N-GPU line's value by N x this for strong-scaling efficiency"}
            c2.close()
        except Exception as e:
            scaling_point = {"workload": "random32", "value": None, "note": f"failed: {e}"}

    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        try:
            gps_step = args.cpu_gates_per_step or (2 if n >= 30 else 4 if n >= 28 else 50)
            gps, sec_step, desc, cores, n_cpu = run_cpu(name, 3, 1, gps_step)
            cpu_baseline = {"value": gps, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as e:  # the CPU leg must never break the GPU number
            cpu_baseline = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": "gates_per_sec", "value": dense_state["value"], "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dense_state["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {workload_desc(name)}", "apply_method": method,
                       "step": "forward of the whole circuit on a DENSE state (no reset between steps, nothing known to be zero: "
                               "every pass reads and writes every amplitude).  `from_reset` is the same circuit as the "
                               "reference's API runs it, reset to |0..0> + forward, where the engine tracks which qubits have "
                               "left |0> and neither reads nor launches tiles that are zero by construction",
                       "l2": "state (>= 4 GiB per GPU) is far larger than the 126 MB L2; no flush needed",
                       "fused": not args.unfused, "wall_s_timed_region": t_wall_dense, "wall_s_from_reset_region": t_wall,
                       "jit": jit_cfg},
            "gpu_launches": int(round(st1["kernel_launches"] * dense_reps)),
            "passes_per_circuit": int(round(st1["tile_passes"])),
            "global_swaps_per_circuit": int(round(st1["global_swaps"])),
            "from_reset": {"value": value, "unit": "gates/s", "ms_per_step": ms / args.steps, "steps": args.steps,
                           "gpu_launches": int(st["kernel_launches"]),
                           "what": "reset to |0..0> + forward, timed like `value` (the step every BASELINE config describes)"},
            "dense_state": dense_state,
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline,
        }
        if scaling_point is not None:
            line["scaling_point"] = scaling_point
        print(json.dumps(line), flush=True)
    if circ is not None:
        circ.close()
    if args.gpus > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
