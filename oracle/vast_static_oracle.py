"""TEST INFRASTRUCTURE ONLY: loop-for-loop restatement of the reference's static design
``Matlab/ControlMethods/vast.m:42-98`` (impulse-driven statistics, joint diagonalisation, rank-V filter sum) and of
``jdiag.m:91-116`` (Cholesky of B, no loading), for small sizes.  Parity unpinned by reference outputs: no MATLAB /
Octave in this environment and the reference ships no vectors for this function; the restatement follows the
MATLAB source statement by statement."""
import numpy as np
import scipy.linalg as sla


def vast_oracle(gB, gD, filter_length, modelling_delay, reference_index, number_of_eigenvectors, mu, N=1000):
    """reference_index is 0-based here (MATLAB: 1-based)."""
    M, I, L = gB.shape
    J = filter_length
    n = J * L
    RB = np.zeros((n, n)); RD = np.zeros((n, n)); rB = np.zeros(n)
    x = np.zeros(N); x[0] = 1.0                                     # vast.m:49
    xPad = np.concatenate([np.zeros(I - 1), x])                     # :50-51
    X = np.zeros((J, I))                                            # :52
    for nIdx in range(N):                                           # :53
        xTmp = xPad[nIdx:nIdx + I][::-1]                            # :54
        X[1:J, :] = X[0:J - 1, :].copy()                            # :55
        X[0, :] = xTmp                                              # :56
        for mIdx in range(M):                                       # :57
            d = X @ np.concatenate([np.zeros(modelling_delay), gB[mIdx, :I - modelling_delay, reference_index]])  # :59
            yB = np.zeros(n); yD = np.zeros(n)
            for s in range(L):                                      # :62-66
                yB[s * J:(s + 1) * J] = X @ gB[mIdx, :, s]
                yD[s * J:(s + 1) * J] = X @ gD[mIdx, :, s]
            RB += np.outer(yB, yB)                                  # :67
            RD += np.outer(yD, yD)                                  # :69
            rB += yB * d[0]                                         # :70
    c = M * (I - J)                                                 # :73-75
    RB /= c; RD /= c; rB /= c
    Bc = np.linalg.cholesky(RD)                                     # jdiag.m:98
    C = sla.solve_triangular(Bc, sla.solve_triangular(Bc, RB, lower=True).T, lower=True).T     # :105
    C = 0.5 * (C + C.T)
    lam, Q = np.linalg.eigh(C)                                      # :106 (schur of a symmetric matrix)
    order = np.argsort(lam)[::-1]                                   # :109
    lam = lam[order]
    U = sla.solve_triangular(Bc.T, Q, lower=False)[:, order]        # :107,111
    w = np.zeros(n)
    for i in range(number_of_eigenvectors):                         # vast.m:87-89
        w = w + (U[:, i] @ rB) / (lam[i] + mu) * U[:, i]
    return w.reshape(L, J).T, RB, RD, rB
