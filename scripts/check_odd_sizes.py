import sys; sys.path.insert(0, "/root/repo")
import numpy as np, time
from ap_vast_unofficial_b200 import jdiag
def spd_pair(n, rng, cols=3):
    X = rng.standard_normal((n, cols * n)); Y = rng.standard_normal((n, cols * n))
    sc = np.exp(-np.arange(n) / (n / 6.0))
    return (X * sc[:, None]) @ (X * sc[:, None]).T, Y @ Y.T
for n in (1025, 1537, 2050, 3001):
    rng = np.random.default_rng(n)
    A, B = spd_pair(n, rng)
    V = 40
    t0 = time.time(); U3, D3 = jdiag(A, B, number_of_eigenvectors=V, eig_mode=3); t3 = time.time() - t0
    U1, D1 = jdiag(A, B, number_of_eigenvectors=V, eig_mode=1)
    l1, l3 = np.diag(D1), np.diag(D3)
    r = rng.standard_normal(n)
    w1 = np.cumsum(U1 * ((U1.T @ r) / (l1 + 0.7))[None, :], axis=1)
    w3 = np.cumsum(U3 * ((U3.T @ r) / (l3 + 0.7))[None, :], axis=1)
    err = np.linalg.norm(w1 - w3, axis=0) / np.linalg.norm(w1, axis=0)
    Breg = B + 1e-7 * np.eye(n)
    print(n, "lam %.2e" % (np.max(np.abs(l1 - l3)) / l1[0]), "filters %.2e" % err.max(), "U'BU-I %.2e" % np.max(np.abs(U3.T @ Breg @ U3 - np.eye(V))), "%.2fs" % t3, flush=True)
