"""time device eigh/SVD (through the C ABI, host buffers) against host LAPACK (numpy) for block sizes seen in DMRG"""
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
print("method", os.environ.get("ITB_SVD_METHOD", "1"), "threads", os.environ.get("OPENBLAS_NUM_THREADS"))
for n in (64, 128, 256, 512, 1024, 1600):
    a = rng.standard_normal((n, n)); a = a + a.T
    w = np.zeros(n); info = C.c_int32()
    def dev():
        A = np.asfortranarray(a.copy())
        check(lib().itb_syevd_host(ctx.handle, 0, n, A.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(info)))
    m2 = 2 * n
    b = rng.standard_normal((n, m2))
    s = np.zeros(n); U = np.zeros((n, n), order="F"); VT = np.zeros((n, m2), order="F")
    def devsvd():
        B = np.asfortranarray(b.copy())
        check(lib().itb_gesvd_host(ctx.handle, 0, n, m2, B.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data_as(C.c_void_p), VT.ctypes.data_as(C.c_void_p), C.byref(info)))
    print(f"n={n:5d}  eigh dev {t(dev):8.2f} ms  host {t(lambda: np.linalg.eigh(a)):8.2f} ms | svd {n}x{m2} dev {t(devsvd):8.2f} ms  host {t(lambda: np.linalg.svd(b, full_matrices=False)):8.2f} ms")
