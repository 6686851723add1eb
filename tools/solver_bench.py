#!/usr/bin/env python
"""Block SVD routes on one B200 (svdBond's per-block factorisation): cuSOLVER polar SVD behind itb_svd_batch_*
against Gram matrix + device eigh (itb_eigh_batch_*) + two GEMMs, for DMRG-like spectra (exponentially decaying singular
values). Prints milliseconds and the accuracy of the Gram route for the leading half of the spectrum."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itensor_b200 as itb
from itensor_b200._lib import check, lib

ctx = itb.Context(0)
dev = ctx.device
I64, I32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
def sync(): torch.cuda.synchronize()
for n in (128, 256, 512, 768, 1024, 1268, 2000):
    g = torch.Generator(device=dev); g.manual_seed(n)
    q1, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    q2, _ = torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, device=dev, generator=g))
    s_true = torch.exp(-torch.arange(n, dtype=torch.float64, device=dev) * (30.0 / n))     # 1 ... 1e-13
    A = (q1 * s_true) @ q2.T                                                              # row-major n x n == column-major A^T: fine, square
    A = A.contiguous()
    off = np.zeros(1, np.int64); mm = np.array([n], np.int32)
    def polar():
        h = C.c_void_p()
        check(lib().itb_svd_batch_run(ctx.handle, 0, 1, off.ctypes.data_as(I64), mm.ctypes.data_as(I32), mm.ctypes.data_as(I32), C.c_void_p(A.data_ptr()), C.byref(h)))
        s = np.zeros(n); check(lib().itb_svd_batch_values(h, s.ctypes.data_as(C.POINTER(C.c_double))))
        check(lib().itb_svd_batch_destroy(h)); return s
    def gram():
        G = A @ A.T                                                                       # cuBLAS DGEMM (the plugin would use its own tile kernel)
        h = C.c_void_p()
        check(lib().itb_eigh_batch_run(ctx.handle, 0, 1, off.ctypes.data_as(I64), mm.ctypes.data_as(I32), C.c_void_p(G.data_ptr()), 1, C.byref(h)))
        w = np.zeros(n); check(lib().itb_eigh_batch_values(h, w.ctypes.data_as(C.POINTER(C.c_double))))
        k = n // 2
        U = torch.zeros(n * k, dtype=torch.float64, device=dev)
        check(lib().itb_eigh_batch_copy_vectors(h, 0, k, C.c_void_p(U.data_ptr()), 0))
        Um = U.view(k, n)                                                                 # column-major n x k == row-major k x n
        V = Um @ A                                                                        # (U^T A): k x n, rows = sigma_i v_i^T
        check(lib().itb_eigh_batch_destroy(h)); sync()
        return np.sqrt(np.maximum(-w, 0)), Um, V
    for f in (polar, gram): f(); sync()
    t = {}
    for name, f in (("polar", polar), ("gram", gram)):
        best = 1e9
        for _ in range(3):
            sync(); t0 = time.perf_counter(); r = f(); sync(); best = min(best, time.perf_counter() - t0)
        t[name] = best * 1e3
    sp = polar(); sg, Um, V = gram()
    k = n // 2
    st = s_true.cpu().numpy()
    err_p = np.abs(sp[:k] - st[:k]).max(); err_g = np.abs(sg[:k] - st[:k]).max()
    rel_g = (np.abs(sg[:k] - st[:k]) / st[:k]).max()
    vn = V / torch.from_numpy(sg[:k]).to(dev)[:, None]
    orth = float((vn @ vn.T - torch.eye(k, dtype=torch.float64, device=dev)).abs().max().item())
    print(f"n={n:5d}: polar SVD {t['polar']:8.2f} ms   Gram+eigh+GEMMs {t['gram']:8.2f} ms   |  leading half (sigma >= {st[k-1]:.1e}): max abs err polar {err_p:.1e} gram {err_g:.1e} (rel {rel_g:.1e}), V orthonormality defect {orth:.1e}", flush=True)
