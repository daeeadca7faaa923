"""NumPy prototype of the divide-and-conquer eigensolver for the symmetric tridiagonal matrix of the two-stage
reduction (north_star (3); Cuppen's method with Gu-Eisenstat vector recomputation).  It has the SAME structure as
csrc/dc.cu -- level-by-level merges, rank sort by counting, deflation scan, one secular root per warp by the bracketed
"middle way" iteration, Loewner recomputation of z, dense merge matrix U~ and one GEMM per half -- so that every
kernel can be checked against its restatement here.

    python scripts/proto_dc.py            # self-test against scipy.linalg.eigh_tridiagonal
"""
import numpy as np

EPS = np.finfo(np.float64).eps
LEAF = 32


def split_diagonal(d, e, leaf=LEAF):
    """T = blockdiag(leaves) + sum_k rho_k u_k u_k^T, u_k = e_{k-1} + sign(e_{k-1}) e_k at every multiple k of `leaf`."""
    d = d.copy()
    n = d.size
    for k in range(leaf, n, leaf):
        r = abs(e[k - 1])
        d[k - 1] -= r
        d[k] -= r
    return d


def secular_root(dd, w, i, rho_total):
    """Root i of f(x) = 1 + sum_j w_j / (dd_j - x) in (dd_i, dd_{i+1}) (last: (dd_{m-1}, dd_{m-1} + rho_total]).
    Returns (K, tau): x = dd_K + tau with tau accurate to a few ulps (so that dd_j - x = (dd_j - dd_K) - tau is accurate).
    Bracketed "middle way" iteration (Li 1994 / LAPACK dlaed4): the poles left of the root (psi) and right of it (phi)
    are each replaced by one pole at the interval end, matched in value and derivative; the two-pole equation is a
    quadratic.  Same arithmetic as dc_secular_kernel."""
    m = dd.size
    last = i == m - 1
    if last:
        K = i
        lo, hi = 0.0, rho_total
    else:
        gap = dd[i + 1] - dd[i]
        mid = 0.5 * gap
        fm = 1.0 + np.sum(w / ((dd - dd[i]) - mid))
        if fm >= 0.0:
            K = i
            lo, hi = 0.0, mid
        else:
            K = i + 1
            lo, hi = -mid, 0.0
    delta = dd - dd[K]
    left = np.arange(m) <= i
    t = 0.5 * (lo + hi)
    for it in range(400):
        q = 1.0 / (delta - t)
        wq = w * q
        psi, dpsi = np.sum(wq[left]), np.sum(wq[left] * q[left])
        phi, dphi = np.sum(wq[~left]), np.sum(wq[~left] * q[~left])
        f = 1.0 + psi + phi
        if f == 0.0:
            break
        if f > 0.0:
            hi = t
        else:
            lo = t
        dl = delta[i] - t
        qq, pp = dpsi * dl * dl, psi - dpsi * dl
        if last:
            c = 1.0 + pp
            tn = delta[i] + qq / c if c != 0.0 else 0.5 * (lo + hi)
        else:
            dr = delta[i + 1] - t
            ss, rr = dphi * dr * dr, phi - dphi * dr
            c = 1.0 + pp + rr
            qa, qb, qc = c, -(c * (dl + dr) + qq + ss), c * dl * dr + qq * dr + ss * dl
            if qa == 0.0:
                eta = -qc / qb if qb != 0.0 else 0.0
            else:
                sq = np.sqrt(max(qb * qb - 4.0 * qa * qc, 0.0))
                e1 = (-qb + sq) / (2.0 * qa) if qb <= 0.0 else (-qb - sq) / (2.0 * qa)
                e2 = qc / (qa * e1) if e1 != 0.0 else 0.0
                eta = e1 if dl < e1 < dr else e2
            tn = t + eta
        if not (lo < tn < hi):
            # bisection; geometric while the bracket still spans orders of magnitude (a root very close to its pole)
            if lo == 0.0:
                tn = hi * 0.015625
            elif hi == 0.0:
                tn = lo * 0.015625
            elif lo > 0.0 and hi > 4.0 * lo:
                tn = np.sqrt(lo * hi)
            elif hi < 0.0 and lo < 4.0 * hi:
                tn = -np.sqrt(lo * hi)
            else:
                tn = 0.5 * (lo + hi)
        if tn == t or abs(tn - t) <= 2.0 * EPS * abs(tn):
            t = tn
            break
        t = tn
        if hi - lo <= 2.0 * EPS * max(abs(lo), abs(hi)):
            break
    return K, t


def merge(d1, d2, Q1, Q2, rho, sgn):
    """Eigen-decomposition of blockdiag(Q1 D1 Q1^T, Q2 D2 Q2^T) + rho u u^T, u = e_last(1) + sgn e_first(2)."""
    k1, k2 = d1.size, d2.size
    k = k1 + k2
    d = np.concatenate([d1, d2])
    z = np.concatenate([Q1[-1, :], sgn * Q2[0, :]])
    Q = np.zeros((k, k))
    Q[:k1, :k1] = Q1
    Q[k1:, k1:] = Q2
    nz_ = np.linalg.norm(z)
    z = z / nz_
    rho = rho * nz_ * nz_
    # rank sort by counting (ties by index)
    idx = np.arange(k)
    rank = np.array([np.sum((d < d[j]) | ((d == d[j]) & (idx < j))) for j in range(k)])
    perm = np.empty(k, int)
    perm[rank] = idx                 # perm[s] = original column of sorted position s
    ds, zs = d[perm].copy(), z[perm].copy()
    tol = 8.0 * EPS * max(np.max(np.abs(ds)), np.max(np.abs(zs)))
    keep = []                        # sorted positions that stay in the secular problem
    defl = np.zeros(k, bool)
    p = -1
    for j in range(k):
        if rho * abs(zs[j]) <= tol:
            defl[j] = True
            continue
        if p >= 0:
            s_, c_ = zs[p], zs[j]
            tau = np.hypot(c_, s_)
            t = ds[j] - ds[p]
            c_, s_ = c_ / tau, -s_ / tau
            if abs(t * c_ * s_) <= tol:
                zs[j], zs[p] = tau, 0.0
                a, b = Q[:, perm[p]].copy(), Q[:, perm[j]].copy()
                Q[:, perm[p]] = c_ * a + s_ * b
                Q[:, perm[j]] = -s_ * a + c_ * b
                t2 = ds[p] * c_ * c_ + ds[j] * s_ * s_
                ds[j] = ds[p] * s_ * s_ + ds[j] * c_ * c_
                ds[p] = t2
                defl[p] = True
                p = j
                continue
            keep.append(p)
        p = j
    if p >= 0:
        keep.append(p)
    keep = np.array(keep, int)
    m = keep.size
    lam = np.empty(k)
    U = np.zeros((k, k))             # U~[original column][new column]
    col = 0
    for j in np.nonzero(defl)[0]:
        lam[col] = ds[j]
        U[perm[j], col] = 1.0
        col += 1
    if m > 0:
        dd, zz = ds[keep], zs[keep]
        w = rho * zz * zz
        rho_total = np.sum(w)
        Ks, taus = np.empty(m, int), np.empty(m)
        for i in range(m):
            Ks[i], taus[i] = secular_root(dd, w, i, rho_total)
        # Loewner: rho zhat_j^2 = prod_i (lam_i - dd_j) / prod_{i != j} (dd_i - dd_j)
        zhat = np.empty(m)
        for j in range(m):
            num = (dd[Ks] - dd[j]) + taus               # lam_i - dd_j, accurate
            prod = num[j]
            for i in range(m):
                if i != j:
                    prod *= num[i] / (dd[i] - dd[j])
            zhat[j] = np.copysign(np.sqrt(abs(prod) / rho), zz[j])
        for i in range(m):
            den = (dd - dd[Ks[i]]) - taus[i]            # dd_j - lam_i
            u = zhat / den
            u /= np.linalg.norm(u)
            lam[col] = dd[Ks[i]] + taus[i]
            U[perm[keep], col] = u
            col += 1
    return lam, Q @ U


def tql2_leaf(d, e):
    """Leaf solver: implicit QL with Wilkinson shifts (EISPACK tql2), the algorithm of the leaf kernel."""
    n = d.size
    d = d.copy()
    e = np.concatenate([e, [0.0]]).copy()
    Z = np.eye(n)
    for l in range(n):
        it = 0
        while True:
            mm = l
            while mm < n - 1:
                if abs(e[mm]) <= EPS * (abs(d[mm]) + abs(d[mm + 1])):
                    break
                mm += 1
            if mm == l:
                break
            it += 1
            assert it < 60
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[mm] - d[l] + e[l] / (g + np.copysign(r, g))
            s = c = 1.0
            p = 0.0
            i = mm - 1
            under = False
            while i >= l:
                f = s * e[i]
                b = c * e[i]
                r = np.hypot(f, g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p
                    e[mm] = 0.0
                    under = True
                    break
                s = f / r
                c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r
                d[i + 1] = g + p
                g = c * r - b
                zi, zi1 = Z[:, i].copy(), Z[:, i + 1].copy()
                Z[:, i + 1] = s * zi + c * zi1
                Z[:, i] = c * zi - s * zi1
                i -= 1
            if under:
                continue
            d[l] -= p
            e[l] = g
            e[mm] = 0.0
    return d, Z


def dc_eigh(d, e, leaf=LEAF):
    n = d.size
    dmod = split_diagonal(d, e, leaf)
    # leaves
    blocks = []
    for k0 in range(0, n, leaf):
        k1 = min(n, k0 + leaf)
        lam, Z = tql2_leaf(dmod[k0:k1], e[k0:k1 - 1])
        blocks.append((k0, k1, lam, Z))
    size = leaf
    while len(blocks) > 1:
        nxt = []
        for a in range(0, len(blocks), 2):
            if a + 1 == len(blocks):
                nxt.append(blocks[a])
                continue
            (a0, a1, l1, Q1), (b0, b1, l2, Q2) = blocks[a], blocks[a + 1]
            rho = abs(e[a1 - 1])
            sgn = 1.0 if e[a1 - 1] >= 0 else -1.0
            lam, Q = merge(l1, l2, Q1, Q2, rho, sgn)
            nxt.append((a0, b1, lam, Q))
        blocks = nxt
        size *= 2
    _, _, lam, Q = blocks[0]
    order = np.argsort(lam)
    return lam[order], Q[:, order]


def _check(name, d, e):
    from scipy.linalg import eigh_tridiagonal
    lam, Q = dc_eigh(d, e)
    ref = eigh_tridiagonal(d, e, eigvals_only=True)
    n = d.size
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    nrm = max(np.max(np.abs(ref)), 1e-300)
    print("%-28s n=%4d  eig err %.2e  orth %.2e  resid %.2e" % (
        name, n, np.max(np.abs(lam - ref)) / nrm, np.max(np.abs(Q.T @ Q - np.eye(n))),
        np.max(np.abs(T @ Q - Q * lam[None, :])) / nrm))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (64, 96, 256, 300):
        _check("random", rng.standard_normal(n), rng.standard_normal(n - 1))
    n = 256
    _check("1-2-1", 2 * np.ones(n), -np.ones(n - 1))
    m = 10
    _check("wilkinson glued", np.tile(np.abs(np.arange(-m, m + 1)), 6).astype(float),
           np.concatenate([np.r_[np.ones(2 * m), 1e-9] for _ in range(6)])[:-1])
    _check("tiny couplings", rng.standard_normal(256), 1e-12 * rng.standard_normal(255))
    # spectrum like the reduced pencil: a few large eigenvalues, a long tail of tiny ones
    lamt = np.concatenate([np.logspace(1, -3, 40), 1e-9 * rng.random(216)])
    A = rng.standard_normal((256, 256)); Qr, _ = np.linalg.qr(A)
    C = (Qr * lamt) @ Qr.T
    from scipy.linalg import hessenberg
    Hh = hessenberg((C + C.T) / 2)
    _check("pencil-like spectrum", np.diag(Hh).copy(), np.diag(Hh, 1).copy())
