"""Device eigh / SVD entry points (cuSOLVER behind the C ABI) vs NumPy/LAPACK on the host: the relations the
reference's own tests pin (matrix_test.cc: A = U diag(w) U^H, A = U diag(s) VT, orthonormality)."""
import ctypes as C

import numpy as np
import pytest

import itensor_b200 as itb
from itensor_b200._lib import check, lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cplx", [False, True])
def test_syevd(ctx, cplx):
    rng = np.random.default_rng(1)
    for n in (1, 7, 64, 130, 300):
        a = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
        a = (a + a.conj().T) / 2
        A = np.asfortranarray(a.astype(np.complex128 if cplx else np.float64))
        w = np.zeros(n)
        info = C.c_int32(-1)
        check(lib().itb_syevd_host(ctx.handle, 1 if cplx else 0, n, A.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(info)))
        assert info.value == 0
        assert np.all(np.diff(w) >= -1e-12)  # ascending, like dsyev
        assert np.allclose(w, np.linalg.eigvalsh(a), rtol=1e-11, atol=1e-11)
        assert np.allclose(A @ np.diag(w) @ A.conj().T, a, atol=1e-10)
        assert np.allclose(A.conj().T @ A, np.eye(n), atol=1e-11)


@pytest.mark.parametrize("cplx", [False, True])
def test_gesvd_tall_and_wide(ctx, cplx):
    rng = np.random.default_rng(2)
    for m, n in ((5, 5), (40, 17), (17, 40), (200, 96), (96, 200), (1, 9)):
        a = rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if cplx else 0)
        dt = np.complex128 if cplx else np.float64
        A = np.asfortranarray(a.astype(dt))
        l = min(m, n)
        s = np.zeros(l)
        U = np.zeros((m, l), dt, order="F")
        VT = np.zeros((l, n), dt, order="F")
        info = C.c_int32(-1)
        check(lib().itb_gesvd_host(ctx.handle, 1 if cplx else 0, m, n, A.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.POINTER(C.c_double)),
                                   U.ctypes.data_as(C.c_void_p), VT.ctypes.data_as(C.c_void_p), C.byref(info)))
        assert info.value == 0
        assert np.allclose(s, np.linalg.svd(a, compute_uv=False), rtol=1e-11, atol=1e-11)
        assert np.allclose(U @ np.diag(s) @ VT, a, atol=1e-10), (m, n)
        assert np.allclose(U.conj().T @ U, np.eye(l), atol=1e-11)
        assert np.allclose(VT @ VT.conj().T, np.eye(l), atol=1e-11)
