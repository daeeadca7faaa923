"""GPU: unit parity of the building-block kernels through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capi():
    from ap_vast_unofficial_b200 import _capi
    return _capi


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 136, 48), (37, 301, 75), (64, 500, 64), (260, 64, 130),
                                   (300, 32, 130), (150, 17, 64), (129, 40, 33)])   # skinny tiles (BN = 32 / 64)
def test_gemm_dmma(ta, tb, M, N, K):
    capi = _capi()
    rng = np.random.default_rng(M * 7 + N * 3 + K + ta * 2 + tb)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    Cm = rng.standard_normal((M, N))
    want = 1.5 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * Cm
    got = Cm.copy()
    capi.check(capi.lib().apv_util_gemm(M, N, K, ta, tb, 1.5, capi.ptr(A), capi.ptr(B), -0.5, capi.ptr(got)))
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-14


@pytest.mark.parametrize("n", [8, 30, 64, 98, 250, 1600, 2048, 2 * 509])
def test_fft_mixed_radix(n):
    capi = _capi()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    xin = np.ascontiguousarray(np.stack([x.real, x.imag], axis=1))
    for inv in (0, 1):
        out = np.zeros((n, 2))
        capi.check(capi.lib().apv_util_fft(n, inv, capi.ptr(xin), capi.ptr(out)))
        got = out[:, 0] + 1j * out[:, 1]
        want = np.fft.ifft(x) * n if inv else np.fft.fft(x)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-14


def _spd_pair(n, rng, cols=3):
    X = rng.standard_normal((n, cols * n))
    Y = rng.standard_normal((n, cols * n))
    sc = np.exp(-np.arange(n) / (n / 6.0))          # spread the spectrum like real statistics
    return (X * sc[:, None]) @ (X * sc[:, None]).T, Y @ Y.T


@pytest.mark.parametrize("n,V,eig_mode", [(24, 24, 1), (24, 24, 2), (70, 10, 1), (70, 10, 2), (100, 100, 1), (100, 100, 2),
                                          (111, 7, 2), (257, 33, 0), (640, 64, 0),
                                          # eig_mode 3 = two-stage (band reduction + bulge chasing); 24: band only,
                                          # 34: the smallest panel, 65 / 97: partial last blocks
                                          (24, 24, 3), (34, 10, 3), (65, 65, 3), (97, 20, 3), (100, 100, 3),
                                          (257, 33, 3), (640, 64, 3), (1100, 64, 3),
                                          # V = 100 / 150: four / five 32-vector groups in the wavefront back-transformation
                                          (700, 100, 3), (300, 150, 3)])
def test_jdiag_identities_and_filters(n, V, eig_mode):
    """jdiag.m:33-35 identities and the filter sum against the reference route (oracle jdiag)."""
    from ap_vast_unofficial_b200 import jdiag
    from oracle.apvast_oracle import jdiag as jdiag_ref
    rng = np.random.default_rng(n + V)
    A, B = _spd_pair(n, rng)
    U, D = jdiag(A, B, number_of_eigenvectors=V, eig_mode=eig_mode)
    lam = np.diag(D)
    Ur, Dr = jdiag_ref(A, B)
    lr = np.diag(Dr)[:V]
    assert np.max(np.abs(lam - lr)) / lr[0] < 1e-12
    Breg = B + 1e-7 * np.eye(n)
    assert np.max(np.abs(U.T @ Breg @ U - np.eye(V))) < 1e-9
    G = U.T @ A @ U
    assert np.max(np.abs(G - np.diag(lam))) / lam[0] < 1e-9
    r = rng.standard_normal(n)
    mu = 0.7
    w = np.cumsum((U * ((U.T @ r) / (lam + mu))[None, :]), axis=1)
    wr = np.cumsum((Ur[:, :V] * ((Ur[:, :V].T @ r) / (lr + mu))[None, :]), axis=1)
    err = np.linalg.norm(w - wr, axis=0) / np.linalg.norm(wr, axis=0)
    assert err.max() < 1e-8, err.max()


def test_jdiag_not_positive_definite():
    from ap_vast_unofficial_b200 import jdiag
    n = 40
    rng = np.random.default_rng(0)
    A = np.eye(n)
    B = rng.standard_normal((n, n)); B = B + B.T - 5 * np.eye(n)
    with pytest.raises(np.linalg.LinAlgError):
        jdiag(A, B, number_of_eigenvectors=4)


@pytest.mark.parametrize("eig_mode", [1, 2, 3])
def test_jdiag_rank_deficient_bright(eig_mode):
    """Degenerate zero eigenvalues (bright matrix of rank 5): the rank-n filter is still the closed form."""
    from ap_vast_unofficial_b200 import jdiag
    n = 48
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, 5)); A = X @ X.T
    Y = rng.standard_normal((n, 2 * n)); B = Y @ Y.T
    U, D = jdiag(A, B, eig_mode=eig_mode)
    lam = np.diag(D)
    r = rng.standard_normal(n)
    w = U @ ((U.T @ r) / (lam + 0.5))
    want = np.linalg.solve(A + 0.5 * (B + 1e-7 * np.eye(n)), r)
    assert np.linalg.norm(w - want) / np.linalg.norm(want) < 1e-8


def test_dmma_peak_reports():
    capi = _capi()
    tf = C.c_double(0)
    capi.check(capi.lib().apv_bench_dmma_peak(2000, C.byref(tf)))
    print("DMMA peak TFLOP/s:", tf.value)
    assert tf.value > 1.0


def test_dfma_rate_reports():
    """CUDA-core FP64 next to the tensor pipe (decides how much scalar FP64 the latency-bound kernels can afford)."""
    capi = _capi()
    o = (C.c_double * 3)()
    capi.check(capi.lib().apv_bench_dfma(1000, o))
    print("DFMA TFLOP/s %.1f, %.1f cycles per dependent DFMA" % (o[0], o[1]))
    assert o[0] > 1.0 and 2.0 < o[1] < 64.0


@pytest.mark.parametrize("n", [96, 200, 1000])
def test_two_stage_matches_one_stage(n):
    """The two tridiagonalisation routes give the same eigenvalues (to rounding) and the same filters."""
    from ap_vast_unofficial_b200 import jdiag
    rng = np.random.default_rng(n)
    A, B = _spd_pair(n, rng)
    V = min(n, 48)
    U1, D1 = jdiag(A, B, number_of_eigenvectors=V, eig_mode=1)
    U3, D3 = jdiag(A, B, number_of_eigenvectors=V, eig_mode=3)
    l1, l3 = np.diag(D1), np.diag(D3)
    assert np.max(np.abs(l1 - l3)) / l1[0] < 1e-13
    r = rng.standard_normal(n)
    w1 = np.cumsum(U1 * ((U1.T @ r) / (l1 + 0.7))[None, :], axis=1)
    w3 = np.cumsum(U3 * ((U3.T @ r) / (l3 + 0.7))[None, :], axis=1)
    err = np.linalg.norm(w1 - w3, axis=0) / np.linalg.norm(w1, axis=0)
    assert err.max() < 1e-8, err.max()
