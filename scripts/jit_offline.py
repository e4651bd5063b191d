"""Offline look at the structure-specialised kernels of a workload: generate each pass's source in every kernel form
and compile it with nvcc -Xptxas -v (no GPU needed): registers, spills, local memory.
    python scripts/jit_offline.py qft30 [n_local] [forms]
"""
import ctypes
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damavand_b200 import _lib, circuits, gates as pgates  # noqa: E402


class Rec:
    def __init__(self):
        self.g = []

    def add_hadamard_gate(self, q): self.g.append(("Hadamard", q, None, None))
    def add_rotation_x_gate(self, q, t): self.g.append(("RotationX", q, None, t))
    def add_rotation_y_gate(self, q, t): self.g.append(("RotationY", q, None, t))
    def add_rotation_z_gate(self, q, t): self.g.append(("RotationZ", q, None, t))
    def add_cnot_gate(self, c, t): self.g.append(("CNOT", t, c, None))
    def add_pauli_z_gate(self, q, obs): pass


def main():
    name = sys.argv[1]
    n, build = circuits.workload(name)
    n_local = int(sys.argv[2]) if len(sys.argv) > 2 else n
    forms = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2]
    r = Rec()
    build(r)
    arr = (_lib.Gate * len(r.g))()
    for k, (nm, t, c, p) in enumerate(r.g):
        arr[k].target = t
        arr[k].control = -1 if c is None else c
        arr[k].m[:] = pgates.matrix(nm, p)
    L = _lib.load()
    buf = ctypes.create_string_buffer(1 << 22)
    seen = {}
    i = 0
    while True:
        k = L.dvd_jit_debug_source(n, n_local, arr, len(r.g), i, 0, buf, 1 << 22)
        if k <= 0:
            break
        seen.setdefault(buf.value, []).append(i)
        i += 1
    print(f"{name}: {i} passes, {len(seen)} distinct structures")
    jobs = []
    for si, (src0, idxs) in enumerate(seen.items()):
        for f in forms:
            L.dvd_jit_debug_source(n, n_local, arr, len(r.g), idxs[0], f, buf, 1 << 22)
            path = f"/tmp/jit/{name}_s{si}_f{f}.cu"
            open(path, "wb").write(buf.value)
            jobs.append((si, f, len(idxs), path))

    def comp(job):
        si, f, cnt, path = job
        res = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-Xptxas", "-v", "-cubin",
                              "-I", os.path.join(ROOT, "damavand_b200", "csrc"), "-o", path.replace(".cu", ".cubin"), path],
                             capture_output=True, text=True)
        out = res.stdout + res.stderr
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", out)
        m2 = re.search(r"Used (\d+) registers", out)
        return si, f, cnt, (m.groups() if m else None), (m2.group(1) if m2 else out[-400:])

    with ThreadPoolExecutor(8) as ex:
        for si, f, cnt, sp, regs in ex.map(comp, jobs):
            print(f"  structure {si:3d} (x{cnt:3d}) form {f}: regs {regs} stack/spill_st/spill_ld {sp}")


if __name__ == "__main__":
    main()
