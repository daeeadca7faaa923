"""Joint-diagonalisation time by eigen-solver path and size (CUDA events through apv_stage_times):
eig_mode 1 = tridiagonalisation + bisection + inverse iteration, eig_mode 2 = shared-memory one-sided Jacobi."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast

def run(L, J, V, mode, M=4, K=64, Nb=256, N=512):
    rng = np.random.default_rng(1)
    dec = np.exp(-np.arange(K) / 16.0).reshape(-1, 1, 1)
    rA = 1e-3 * rng.standard_normal((K, L, M)) * dec; rB = 1e-3 * rng.standard_normal((K, L, M)) * dec
    np.random.seed(0)
    e = apvast(Nb, rA, rB, J, 2, 0, 0, V, 1.0, N, perceptual=False, eig_mode=mode)
    ts = []
    for t in range(6):
        e.process_input_buffers(rng.standard_normal(Nb // 2), rng.standard_normal(Nb // 2))
        st = e.stage_times()
        ts.append((st["S5_jdiag"], st["S5_tridiag"] + st["S5_eig"] + st["S5_backtransform"]))
    e.close()
    return np.median([a for a, _ in ts]), np.median([b for _, b in ts])

print("%6s %4s | %22s | %22s" % ("n", "V", "tridiag path: jdiag / eig ms", "Jacobi path: jdiag / eig ms"))
for L, J in ((2, 8), (4, 8), (4, 12), (4, 16), (4, 24), (4, 28)):
    n = L * J
    for V in (min(8, n), n):
        a = run(L, J, V, 1); b = run(L, J, V, 2)
        print("%6d %4d | %10.3f / %8.3f | %10.3f / %8.3f" % (n, V, a[0], a[1], b[0], b[1]))
