"""Passes per step of the distributed schedule for a workload (host only: the planner through the CPU replay library)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damavand_b200 import circuits
from oracle.oracle import OracleCircuit
from tests.helpers import gate_array


def load():
    """DVD_STORE_REMAP=0/1/2 selects the store-side mode (default 2), DVD_DEFER_TAILS=0 turns tail deferral off."""
    so = os.path.join(ROOT, "tests", "emu", "libdvd_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    deps = [src] + [os.path.join(ROOT, "damavand_b200", "csrc", f) for f in ("planner.cpp", "planner.h", "tile_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emu_plan_only.restype = ctypes.c_int64
    L.emu_error.restype = ctypes.c_char_p
    L.emu_set_store(int(os.environ.get("DVD_STORE_REMAP", "2")))
    return L


def stats(L, name, world):
    n, build = circuits.workload(name)
    o = OracleCircuit.__new__(OracleCircuit); o.num_qubits = n; o.gates = []; o.observables = []
    build(o)
    arr, ng = gate_array(o)
    cap = 1 << 16
    out = (ctypes.c_int32 * cap)()
    k = L.emu_plan_only(n, world, arr, ng, out, cap)
    assert k > 0, L.emu_error()
    pos = 1; desc = []; total = 0; swaps = 0
    for _ in range(out[0]):
        kind, a, b, ngt, npass = out[pos:pos + 5]; pos += 5
        if kind == 0:
            ops = list(out[pos:pos + npass]); pos += npass
            desc.append(f"L{ngt}:{ops}")
            total += npass
        else:
            # S = swap riding on the next load, X = local transposition (restore); P / PX = the same riding on the store of the pass in front
            desc.append(f"{ {1: 'S', 2: 'X', 3: 'P', 4: 'PX'}[kind]}({a},{b})"); swaps += kind in (1, 3)
    print(f"{name} world={world}: gates {ng}, passes {total}, swaps {swaps}")
    print("   " + " ".join(desc))
    return total


if __name__ == "__main__":
    L = load()
    for w in [int(x) for x in sys.argv[2:]] or [8]:
        stats(L, sys.argv[1], w)
