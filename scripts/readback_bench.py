"""Host <-> device transfer rate of dvd_read_state / dvd_load_state (separate real / imaginary arrays, pageable host memory,
as the reference's retrieve_amplitudes_on_host / load_amplitudes_local_on_device take them)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from damavand_b200 import Circuit, _lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    c = Circuit(n, "gpu")
    for q in range(n):
        c.add_hadamard_gate(q)
    c.forward()
    N = 1 << n
    re = np.empty(N); im = np.empty(N)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    L = c._lib
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(L.dvd_read_state(c._handle, dp(re), dp(im), 0, N), "read")
        t1 = time.perf_counter()
        _lib.check(L.dvd_load_state(c._handle, dp(re), dp(im), 0, N), "load")
        t2 = time.perf_counter()
        print(f"n={n} ({16 * N / 2**30:.0f} GiB) rep {rep}: read_state {t1 - t0:.3f} s = {16 * N / (t1 - t0) / 1e9:.2f} GB/s, "
              f"load_state {t2 - t1:.3f} s = {16 * N / (t2 - t1) / 1e9:.2f} GB/s", flush=True)
    assert abs(re[0] - 2.0 ** (-n / 2)) < 1e-15 and abs(c.norm() - 1.0) < 1e-12
    print("values ok")


if __name__ == "__main__":
    main()
