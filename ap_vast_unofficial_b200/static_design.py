"""Static VAST filter design from room impulse responses (reference
``Matlab/ControlMethods/vast.m``): the statistics are those of an impulse driving the RIRs, i.e. correlations of
the impulse responses themselves (a degenerate case of S4), followed by the same joint diagonalisation and
rank-V filter sum as the block engine (S5, S6).

The statistics are assembled on the host (setup-time work, one GEMM per microphone); the joint diagonalisation runs
on the GPU through the C-ABI (``apv_jdiag``).
"""
from __future__ import annotations

import numpy as np


def static_statistics(gB, gD, filter_length: int, modelling_delay: int, reference_index: int, n_samples: int = 1000):
    """R_B, R_D (n x n) and r_B (n,) of vast.m:42-75, n = filter_length * L.

    gB, gD: (M, I, L) impulse responses microphone x tap x loudspeaker (vast.m:9-19).  reference_index is 0-based.
    With an impulse input the delay-line matrix X of vast.m:50-57 gives y[(s, i)] = g[m, t - i, s] at time t, so per
    microphone the statistics are Y^T Y with Y[t, (s, i)] = g[m, t - i, s], t = 0 .. n_samples - 1 (vast.m fixes
    n_samples = 1000, which truncates responses longer than that exactly as the reference does)."""
    gB = np.asarray(gB, dtype=np.float64)
    gD = np.asarray(gD, dtype=np.float64)
    M, I, L = gB.shape
    J, N = int(filter_length), int(n_samples)
    n = J * L
    t = np.arange(N)[:, None] - np.arange(J)[None, :]            # tap index t - i
    valid = (t >= 0) & (t < I)
    tc = np.clip(t, 0, I - 1)

    def delay_matrix(g_m):                                       # (I, L) -> (N, L * J), loudspeaker-major columns
        Y = np.where(valid[:, None, :], g_m[tc].transpose(0, 2, 1), 0.0)      # Y[t, s, i] = g_m[t - i, s]
        return Y.reshape(N, n)

    RB = np.zeros((n, n)); RD = np.zeros((n, n)); rB = np.zeros(n)
    for m in range(M):
        YB = delay_matrix(gB[m])
        d = np.zeros(N)                                          # target: the reference loudspeaker's RIR, delayed
        tgt = np.concatenate([np.zeros(modelling_delay), gB[m, :I - modelling_delay, reference_index]])
        d[:min(N, I)] = tgt[:min(N, I)]
        RB += YB.T @ YB
        rB += YB.T @ d
    for m in range(gD.shape[0]):                                 # (the zones may have different microphone counts)
        YD = delay_matrix(gD[m])
        RD += YD.T @ YD
    c = float(M * (I - J))
    return RB / c, RD / c, rB / c


def vast_static(gB, gD, filter_length: int, modelling_delay: int, reference_index: int, number_of_eigenvectors: int,
                mu: float, n_samples: int = 1000, reg: float = 0.0, eig_mode: int = 0):
    """w (filter_length, L) of vast.m:78-98: w = sum_{v < V} (u_v^T r_B) / (lambda_v + mu) u_v with (U, lambda) the joint
    diagonalisation of (R_B, R_D).  ``reg`` = 0 is the MATLAB jdiag (no loading; numpy.linalg.LinAlgError when R_D
    is not positive definite, like its ``error``)."""
    from .apvast import jdiag

    RB, RD, rB = static_statistics(gB, gD, filter_length, modelling_delay, reference_index, n_samples)
    V = int(number_of_eigenvectors)
    U, D = jdiag(RB, RD, number_of_eigenvectors=V, reg=reg, eig_mode=eig_mode)
    lam = np.diag(D)
    w = U @ ((U.T @ rB) / (lam + mu))
    L = np.asarray(gB).shape[2]
    return w.reshape(L, filter_length).T.copy()
