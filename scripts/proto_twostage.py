"""NumPy prototype of the two-stage tridiagonalisation implemented in csrc/band.cu (design aid, not product code):
stage 1  dense -> band (panel QR + two-sided block update), stage 2 band -> tridiagonal by bulge chasing with the
same sweep/step decomposition and reflector storage as the CUDA kernels, then the two back-transformations.
Run:  python scripts/proto_twostage.py [n] [b]"""
import sys
import numpy as np


def house(x):
    """LAPACK dlarfg: H x = beta e_0, H = I - tau v v^T, v[0] = 1."""
    alpha = x[0]
    ss = float(np.dot(x[1:], x[1:]))
    if ss == 0.0:
        v = np.zeros_like(x); v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, np.sqrt(ss)), alpha)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta); v[0] = 1.0
    return v, tau, beta


def stage1(A, b):
    """returns band matrix (dense storage), VH (row j = reflector j, support i >= j + b), tau."""
    n = A.shape[0]
    A = A.copy()
    VH = np.zeros((n, n)); tau = np.zeros(n)
    j0 = 0
    while n - j0 - b >= 2:
        r = j0 + b
        P = A[r:, j0:j0 + b].copy()
        npn, pw = P.shape
        V = np.zeros((npn, pw)); taus = np.zeros(pw)
        for c in range(min(pw, npn - 1)):
            # one Gram row per column: s = P[c:, c:]^T P[c:, c]
            s = P[c:, c:].T @ P[c:, c]
            alpha = P[c, c]
            ss = s[0] - alpha * alpha
            if ss <= 0.0:
                V[c, c] = 1.0
                continue
            beta = -np.copysign(np.sqrt(s[0]), alpha)
            t = (beta - alpha) / beta
            v = P[c:, c] / (alpha - beta); v[0] = 1.0
            w = (s - beta * P[c, c:]) / (alpha - beta)      # v^T P[c:, c:]
            P[c:, c:] -= t * np.outer(v, w)
            P[c + 1:, c] = 0.0
            P[c, c] = beta
            V[c:, c] = v; taus[c] = t
        # T factor (dlarft forward columnwise)
        G = V.T @ V
        T = np.zeros((pw, pw))
        for i in range(pw):
            T[i, i] = taus[i]
            T[:i, i] = -taus[i] * (T[:i, :i] @ G[:i, i])
        A[r:, j0:j0 + b] = P
        A[j0:j0 + b, r:] = P.T
        A22 = A[r:, r:]
        Y = A22 @ V
        S = V.T @ Y
        W = Y @ T - 0.5 * V @ (T.T @ S @ T)
        A22 -= W @ V.T + V @ W.T
        for c in range(pw):
            VH[j0 + c, r:] = V[:, c]; tau[j0 + c] = taus[c]
        j0 += b
    return A, VH, tau


def chase(Aband, b):
    """stage 2 on dense storage.  Returns d, e and V2 (row s = reflectors of sweep s laid out by the rows they act on;
    the first entry of every reflector holds its tau, the implicit leading 1 is not stored)."""
    n = Aband.shape[0]
    A = Aband.copy()
    V2 = np.zeros((n, n))

    def apply_two_sided(D, v, t):
        # D <- H D H, symmetric
        p = t * (D @ v)
        w = p - 0.5 * t * np.dot(p, v) * v
        D -= np.outer(v, w) + np.outer(w, v)

    for s in range(n - 2):
        r0, r1 = s + 1, min(s + 1 + b, n)
        if r1 - r0 < 2:
            break
        v, t, beta = house(A[r0:r1, s].copy())
        A[r0:r1, s] = 0.0; A[r0, s] = beta
        A[s, r0:r1] = A[r0:r1, s]
        apply_two_sided(A[r0:r1, r0:r1], v, t)
        V2[s, r0:r1] = v; V2[s, r0] = t
        while True:
            q0, q1 = r1, min(r1 + b, n)
            if q0 >= n:
                break
            B = A[q0:q1, r0:r1]
            B -= t * np.outer(B @ v, v)                 # B <- B H
            if q1 - q0 >= 2:
                v2, t2, beta = house(B[:, 0].copy())
                B[:, 0] = 0.0; B[0, 0] = beta
                B[:, 1:] -= t2 * np.outer(v2, v2 @ B[:, 1:])     # rest of B <- H' B
            else:
                v2, t2 = np.ones(1), 0.0
            A[r0:r1, q0:q1] = B.T
            if t2 != 0.0:
                apply_two_sided(A[q0:q1, q0:q1], v2, t2)
            V2[s, q0:q1] = v2; V2[s, q0] = t2
            r0, r1, v, t = q0, q1, v2, t2
    return np.diag(A).copy(), np.diag(A, -1).copy(), V2, A


def apply_q2(V2, Z, b):
    """Z <- Q2 Z: sweeps in descending order, the reflectors of a sweep commute."""
    n = V2.shape[0]
    Z = Z.copy()
    for s in range(n - 3, -1, -1):
        r0 = s + 1
        while r0 < n:
            r1 = min(r0 + b, n)
            t = V2[s, r0]
            if t != 0.0:
                v = V2[s, r0:r1].copy(); v[0] = 1.0
                Z[r0:r1] -= t * np.outer(v, v @ Z[r0:r1])
            r0 = r1
    return Z


def apply_q1(VH, tau, Z):
    n = VH.shape[0]
    Z = Z.copy()
    for j in range(n - 1, -1, -1):
        if tau[j] != 0.0:
            Z -= tau[j] * np.outer(VH[j], VH[j] @ Z)
    return Z


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 97
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, n)); A = X + X.T
    Ab, VH, tau = stage1(A, b)
    i, j = np.indices((n, n))
    print("outside band:", np.abs(Ab[np.abs(i - j) > b]).max())
    print("eig band vs A:", np.abs(np.linalg.eigvalsh(Ab) - np.linalg.eigvalsh(A)).max())
    d, e, V2, At = chase(Ab, b)
    print("outside tridiagonal:", np.abs(At[np.abs(i - j) > 1]).max())
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, Z = np.linalg.eigh(T)
    print("eig T vs A:", np.abs(lam - np.linalg.eigvalsh(A)).max())
    Q = apply_q1(VH, tau, apply_q2(V2, Z, b))
    print("residual A Q - Q lam:", np.abs(A @ Q - Q * lam).max(), " orth:", np.abs(Q.T @ Q - np.eye(n)).max())
