"""SVD variants on DMRG-like blocks (decaying spectrum): time and accuracy vs host LAPACK gesdd (numpy).
ITB_SVD_METHOD=0|1|2 python tools/solver_bench3.py"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itensor_b200 as itb
from itensor_b200._lib import check, lib
ctx = itb.Context(0)
rng = np.random.default_rng(0)
print("method", os.environ.get("ITB_SVD_METHOD", "1"), "host threads", os.environ.get("OPENBLAS_NUM_THREADS"))
for n in (256, 512, 1024, 1268, 2048):
    # DMRG-like: orthogonal factors, singular values decaying over 12 decades
    q1, _ = np.linalg.qr(rng.standard_normal((n, n))); q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.exp(-np.arange(n) * (27.6 / n))
    a = (q1 * sv) @ q2.T
    s = np.zeros(n); U = np.zeros((n, n), order="F"); VT = np.zeros((n, n), order="F"); info = C.c_int32()
    def dev():
        B = np.asfortranarray(a.copy())
        check(lib().itb_gesvd_host(ctx.handle, 0, n, n, B.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.POINTER(C.c_double)), U.ctypes.data_as(C.c_void_p), VT.ctypes.data_as(C.c_void_p), C.byref(info)))
    dev()
    t0 = time.perf_counter(); dev(); td = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); uh, sh, vh = np.linalg.svd(a, full_matrices=False); th = (time.perf_counter() - t0) * 1e3
    rec = np.abs((U * s) @ VT - a).max()
    orthU = np.abs(U.T @ U - np.eye(n)).max(); orthV = np.abs(VT @ VT.T - np.eye(n)).max()
    rel = np.abs(s - sh) / sh[0]
    print(f"n={n:5d} dev {td:8.1f} ms host {th:8.1f} ms | recon {rec:.1e} orthU {orthU:.1e} orthV {orthV:.1e} max|ds|/s0 {rel.max():.1e} "
          f"rel err of s at 1e-6,1e-10 levels: {abs(s[n//2]-sh[n//2])/sh[n//2]:.1e} {abs(s[int(n*0.83)]-sh[int(n*0.83)])/sh[int(n*0.83)]:.1e} info {info.value}")
