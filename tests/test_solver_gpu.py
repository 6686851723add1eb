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


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_batch_device_resident(ctx, cplx):
    """itb_svd_batch_*: the blocks of an order-2 block-sparse tensor factorised from its device buffer (svdImpl's QN
    loop, svd.cc:199-222, on QDenseGPU): singular values vs LAPACK, A_b = U_b diag(s_b) V_b^H from the kept leading
    columns copied device-to-device, both solver routes (polar for large blocks, Jacobi for small ones), ragged and
    degenerate (rank-deficient) blocks."""
    import torch

    rng = np.random.default_rng(3)
    dt = np.complex128 if cplx else np.float64
    shapes = [(150, 150), (7, 3), (1, 1), (64, 200), (200, 64), (130, 97), (40, 40), (120, 120)]
    blocks = []
    for i, (m, n) in enumerate(shapes):
        a = rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if cplx else 0)
        if i == 6:
            a[:, 20:] = a[:, :20]  # rank 20 (Jacobi route)
        if i == 7:
            a[:, 30:] = np.tile(a[:, :30], 3)  # rank 30 (polar route)
        blocks.append(np.asfortranarray(a.astype(dt)))
    flat = np.concatenate([b.reshape(-1, order="F") for b in blocks])
    d = torch.from_numpy(flat.view(np.float64)).to(ctx.device)
    before = d.clone()
    off = np.cumsum([0] + [m * n for m, n in shapes[:-1]]).astype(np.int64)
    mm = np.array([s[0] for s in shapes], np.int32)
    nn = np.array([s[1] for s in shapes], np.int32)
    h = C.c_void_p()
    check(lib().itb_svd_batch_run(ctx.handle, 1 if cplx else 0, len(shapes), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                  mm.ctypes.data_as(C.POINTER(C.c_int32)), nn.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(d.data_ptr()), C.byref(h)))
    ls = [min(m, n) for m, n in shapes]
    s = np.zeros(sum(ls))
    check(lib().itb_svd_batch_values(h, s.ctypes.data_as(C.POINTER(C.c_double))))
    assert torch.equal(d, before)  # the input tensor is not modified
    cs = 2 if cplx else 1
    p = 0
    for b, (m, n) in enumerate(shapes):
        l = ls[b]
        sb = s[p:p + l]
        p += l
        ref = np.linalg.svd(blocks[b], compute_uv=False)
        assert np.all(np.diff(sb) <= 1e-12 * ref[0])  # descending
        assert np.allclose(sb, ref, rtol=0, atol=1e-12 * ref[0]), (m, n)
        for k in (l, max(1, l // 2)):  # all columns, then a truncated set
            U = torch.zeros(m * k * cs, dtype=torch.float64, device=ctx.device)
            V = torch.zeros(n * k * cs, dtype=torch.float64, device=ctx.device)
            check(lib().itb_svd_batch_copy_u(h, b, k, C.c_void_p(U.data_ptr())))
            check(lib().itb_svd_batch_copy_v(h, b, k, C.c_void_p(V.data_ptr()), 0))
            torch.cuda.synchronize()
            Uh = U.cpu().numpy().view(dt).reshape(m, k, order="F")
            Vh = V.cpu().numpy().view(dt).reshape(n, k, order="F")
            rec = (Uh * sb[:k]) @ Vh.conj().T
            best = (np.linalg.svd(blocks[b], full_matrices=False)[0][:, :k] * ref[:k]) @ np.linalg.svd(blocks[b], full_matrices=False)[2][:k]
            assert np.abs(rec - best).max() <= 1e-10 * ref[0], (m, n, k)
            nz = int((ref[:k] > 1e-10 * ref[0]).sum())  # singular vectors of the numerical null space are arbitrary
            assert np.allclose(Uh[:, :nz].conj().T @ Uh[:, :nz], np.eye(nz), atol=1e-10)
            assert np.allclose(Vh[:, :nz].conj().T @ Vh[:, :nz], np.eye(nz), atol=1e-10)
        if cplx:  # conj flag
            V0 = torch.zeros(n * l * cs, dtype=torch.float64, device=ctx.device)
            V1 = torch.zeros(n * l * cs, dtype=torch.float64, device=ctx.device)
            check(lib().itb_svd_batch_copy_v(h, b, l, C.c_void_p(V0.data_ptr()), 0))
            check(lib().itb_svd_batch_copy_v(h, b, l, C.c_void_p(V1.data_ptr()), 1))
            torch.cuda.synchronize()
            assert np.array_equal(V1.cpu().numpy().view(dt), V0.cpu().numpy().view(dt).conj())
    check(lib().itb_svd_batch_destroy(h))


@pytest.mark.parametrize("cplx", [False, True])
def test_eigh_batch_device_resident(ctx, cplx):
    """itb_eigh_batch_*: the diagonal blocks of an order-2 block-sparse Hermitian tensor diagonalised from its device
    buffer (diagHImpl's QN loop, hermitian.cc:231-257, on QDenseGPU — the density-matrix branch of svdBond). With
    negate=1 the batch diagonalises -A_b, so -w is the reference's largest-first order (tensor/algs_impl.h:123-137)."""
    import torch

    rng = np.random.default_rng(5)
    dt = np.complex128 if cplx else np.float64
    sizes = [130, 1, 7, 96, 257, 40]
    blocks = []
    for i, n in enumerate(sizes):
        a = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
        a = a @ a.conj().T / n                      # positive semi-definite like a density matrix
        if i == 5:
            a = a[:, :10] @ a[:, :10].conj().T      # rank 10: degenerate zero eigenvalues
        blocks.append(np.asfortranarray(((a + a.conj().T) / 2).astype(dt)))
    flat = np.concatenate([b.reshape(-1, order="F") for b in blocks])
    d = torch.from_numpy(flat.view(np.float64)).to(ctx.device)
    before = d.clone()
    off = np.cumsum([0] + [n * n for n in sizes[:-1]]).astype(np.int64)
    nn = np.array(sizes, np.int32)
    h = C.c_void_p()
    check(lib().itb_eigh_batch_run(ctx.handle, 1 if cplx else 0, len(sizes), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                   nn.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(d.data_ptr()), 1, C.byref(h)))
    w = np.zeros(sum(sizes))
    check(lib().itb_eigh_batch_values(h, w.ctypes.data_as(C.POINTER(C.c_double))))
    assert torch.equal(d, before)  # the input tensor is not modified
    cs = 2 if cplx else 1
    p = 0
    for b, n in enumerate(sizes):
        ev = -w[p:p + n]            # largest first
        p += n
        ref = np.linalg.eigvalsh(blocks[b])[::-1]
        assert np.all(np.diff(ev) <= 1e-13 * max(ref[0], 1e-300))
        assert np.allclose(ev, ref, rtol=0, atol=1e-13 * max(ref[0], 1.0)), n
        for k in (n, max(1, n // 3)):
            U = torch.zeros(n * k * cs, dtype=torch.float64, device=ctx.device)
            check(lib().itb_eigh_batch_copy_vectors(h, b, k, C.c_void_p(U.data_ptr()), 0))
            torch.cuda.synchronize()
            Uh = U.cpu().numpy().view(dt).reshape(n, k, order="F")
            assert np.allclose(Uh.conj().T @ Uh, np.eye(k), atol=1e-11)
            assert np.abs(blocks[b] @ Uh - Uh * ev[:k]).max() <= 1e-11 * max(ref[0], 1.0)   # A u = lambda u
        if cplx:
            U1 = torch.zeros(n * n * cs, dtype=torch.float64, device=ctx.device)
            U0 = torch.zeros(n * n * cs, dtype=torch.float64, device=ctx.device)
            check(lib().itb_eigh_batch_copy_vectors(h, b, n, C.c_void_p(U0.data_ptr()), 0))
            check(lib().itb_eigh_batch_copy_vectors(h, b, n, C.c_void_p(U1.data_ptr()), 1))
            torch.cuda.synchronize()
            assert np.array_equal(U1.cpu().numpy().view(dt), U0.cpu().numpy().view(dt).conj())
    check(lib().itb_eigh_batch_destroy(h))
