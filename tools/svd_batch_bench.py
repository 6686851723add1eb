"""Device-resident batched SVD (itb_svd_batch_run) on the block sizes of one DMRG bond: time vs lanes."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itensor_b200 as itb
from itensor_b200._lib import check, lib
ctx = itb.Context(0)
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "684,631,432,338,200,160").split(",")]
rng = np.random.default_rng(0)
blocks = []
for n in sizes:
    q1, _ = np.linalg.qr(rng.standard_normal((n, n))); q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    blocks.append(np.asfortranarray((q1 * np.exp(-np.arange(n) * (27.6 / n))) @ q2.T))
flat = np.concatenate([b.reshape(-1, order="F") for b in blocks])
d = torch.from_numpy(flat).to(ctx.device)
off = np.cumsum([0] + [n * n for n in sizes[:-1]]).astype(np.int64)
m = np.array(sizes, np.int32)
while not lib().itb_solver_ready(): time.sleep(0.5)
def run():
    h = C.c_void_p()
    check(lib().itb_svd_batch_run(ctx.handle, 0, len(sizes), off.ctypes.data_as(C.POINTER(C.c_int64)), m.ctypes.data_as(C.POINTER(C.c_int32)),
                                  m.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(d.data_ptr()), C.byref(h)))
    s = np.zeros(sum(sizes))
    check(lib().itb_svd_batch_values(h, s.ctypes.data_as(C.POINTER(C.c_double))))
    lib().itb_svd_batch_destroy(h)
    return s
s = run(); run()
t0 = time.perf_counter(); reps = 5
for _ in range(reps): run()
dt = (time.perf_counter() - t0) / reps * 1e3
ref = np.concatenate([np.linalg.svd(b, compute_uv=False) for b in blocks])
print(f"lanes {os.environ.get('ITB_SVD_LANES','4')} sizes {sizes}: {dt:.1f} ms per batch; max |ds|/s0 {np.abs(s-ref).max():.1e}")
for n in sizes:
    single = np.array([n], np.int32); o1 = np.zeros(1, np.int64)
    def one():
        h = C.c_void_p()
        check(lib().itb_svd_batch_run(ctx.handle, 0, 1, o1.ctypes.data_as(C.POINTER(C.c_int64)), single.ctypes.data_as(C.POINTER(C.c_int32)),
                                      single.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(d.data_ptr()), C.byref(h)))
        s1 = np.zeros(n); check(lib().itb_svd_batch_values(h, s1.ctypes.data_as(C.POINTER(C.c_double)))); lib().itb_svd_batch_destroy(h)
    one(); t0 = time.perf_counter(); one(); one(); print(f"   single {n}^2: {(time.perf_counter()-t0)/2*1e3:.1f} ms")
