"""device eigh on DMRG-like density matrices: rank-deficient (rho = X X^T, X n x n/4) and exponentially decaying spectra"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensor_b200 as itb
from itensor_b200._lib import check, lib
ctx = itb.Context(0)
rng = np.random.default_rng(0)
def t(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
for n in (158, 316, 1024):
    for kind in ("random", "rankdef", "decay"):
        if kind == "random":
            a = rng.standard_normal((n, n)); a = a + a.T
        elif kind == "rankdef":
            x = rng.standard_normal((n, n // 4)); a = x @ x.T
        else:
            q, _ = np.linalg.qr(rng.standard_normal((n, n))); a = (q * np.exp(-np.arange(n) * 40.0 / n)) @ q.T; a = (a + a.T) / 2
        w = np.zeros(n); info = C.c_int32()
        def dev():
            A = np.asfortranarray(a.copy())
            check(lib().itb_syevd_host(ctx.handle, 0, n, A.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(info)))
        print(f"n={n:5d} {kind:8s} eigh dev {t(dev):9.2f} ms  host {t(lambda: np.linalg.eigh(a)):8.2f} ms  info {info.value}", flush=True)
