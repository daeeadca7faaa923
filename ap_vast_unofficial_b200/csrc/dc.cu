// Divide-and-conquer eigensolver for the symmetric tridiagonal matrix T of the two-stage reduction, all n eigenpairs
// (north_star (3): "blocked Householder tridiagonalisation plus divide-and-conquer"; the reference forms all n pairs
// with a dense Schur decomposition, Python/apvast.py:30-36, and the rank loop of :406-414 may consume all of them:
// BASELINE cfg-4).  Cuppen's method with Gu-Eisenstat vector recomputation:
//
//   T = blockdiag(leaves of 32) + sum_k rho_k u_k u_k^T      (u_k = e_{k-1} + sign(e_{k-1}) e_k at every multiple of 32)
//   leaves:   implicit QL with Wilkinson shifts (EISPACK tql2), one warp per leaf
//   merges:   level by level, all merges of a level batched; per merge
//             z = Q^T u, rank sort by counting, deflation scan (negligible z; close poles by a Givens rotation),
//             one secular root per warp (bracketed "middle way" iteration, origin at the nearer pole, so that every
//             difference d_j - lambda_i keeps its relative accuracy), Loewner recomputation of z, the dense merge
//             matrix U~ (deflated columns = unit vectors, rotations folded in as row operations), and
//             Q_new = blockdiag(Q_1, Q_2) U~ as ONE batched FP64 tensor-core GEMM per zone (split = the two halves).
//
// The O(n^3) part (4/3 n^3 flops without deflation) runs on the DMMA pipe; everything else is O(n^2).
// scripts/proto_dc.py is the NumPy restatement with the same structure.  Used for full-spectrum requests (V = n) with
// n = 32 * 2^L; other sizes keep multisection + inverse iteration.
#include <float.h>
#include <math.h>

#include <algorithm>

#include "engine.cuh"

namespace apv {

namespace {

constexpr int DC_LEAF = 32;

// per-zone scratch layout (doubles then ints), see dc_scratch_doubles
struct DcWs {
  double* dm;      // [n] eigenvalues of the current level's blocks (any order inside a block)
  double* ds;      // [n] sorted poles of every merge
  double* zs;      // [n] z in sorted order (normalised), zeroed where deflated
  double* dd;      // [n] poles that stay in the secular problem (first m of every merge block)
  double* ww;      // [n] their weights rho z^2
  double* zz;      // [n] their z (sign for the Loewner vector)
  double* taus;    // [n]
  double* zhat;    // [n]
  double* lamn;    // [n] eigenvalues of the merged block, in the column order of U~
  double* rotc;    // [n]
  double* rots;    // [n]
  double* rho;     // [n / 64] scaled rho per merge (+ sum of weights)
  int* perm;       // [n] original column of sorted position
  int* keep;       // [n] sorted positions that stay (first m)
  int* Ks;         // [n]
  int* rotp;       // [n] original columns of the rotated pairs
  int* rotj;       // [n]
  int* dcol;       // [n] sorted positions that are deflated (first nd)
  int* newcol;     // [n] column of U~ (final level: descending rank) of entry c of lamn
  int* meta;       // [n / 64][4] m, ndefl, nrot, unused
};

__host__ __device__ inline size_t dc_doubles_per_zone(int n) { return (size_t)11 * n + n / 64 * 2 + 8; }
__host__ __device__ inline size_t dc_ints_per_zone(int n) { return (size_t)7 * n + (size_t)(n / 64 + 1) * 4; }

__host__ __device__ inline DcWs dc_ws(double* base, int n, int z) {
  DcWs w;
  double* p = base + (size_t)z * (dc_doubles_per_zone(n) + (dc_ints_per_zone(n) + 1) / 2);
  w.dm = p; p += n; w.ds = p; p += n; w.zs = p; p += n; w.dd = p; p += n; w.ww = p; p += n; w.zz = p; p += n;
  w.taus = p; p += n; w.zhat = p; p += n; w.lamn = p; p += n; w.rotc = p; p += n; w.rots = p; p += n;
  w.rho = p; p += n / 64 * 2 + 8;
  int* q = reinterpret_cast<int*>(p);
  w.perm = q; q += n; w.keep = q; q += n; w.Ks = q; q += n; w.rotp = q; q += n; w.rotj = q; q += n; w.dcol = q; q += n;
  w.newcol = q; q += n; w.meta = q;
  return w;
}

// ------------------------------------------------------------------------------------------------
// Leaves.  grid (n / 32 / 4, nz), 128 threads: one warp per 32 x 32 leaf.  The scalar QL recurrences are uniform over
// the warp; lane r owns row r of the eigenvector matrix Z (shared memory, pitch 33).
__global__ void __launch_bounds__(128) dc_leaf_kernel(const double* __restrict__ dd, const double* __restrict__ ee,
                                                      double* __restrict__ scratch, double* __restrict__ Q, long long qstride,
                                                      int ldq, int n, int* __restrict__ info) {
  __shared__ double Zs[4][32][33];
  __shared__ double ds[4][32], es[4][33];
  const int z = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int leaf = blockIdx.x * 4 + w;
  if (leaf * DC_LEAF >= n) return;
  const int k0 = leaf * DC_LEAF;
  const double* gd = dd + (size_t)z * n;
  const double* ge = ee + (size_t)z * n;
  DcWs ws = dc_ws(scratch, n, z);
  // modified diagonal: the couplings to the neighbouring leaves are taken out (rank-one terms of the merges)
  double dv = gd[k0 + lane];
  if (lane == 0 && k0 > 0) dv -= fabs(ge[k0 - 1]);
  if (lane == 31 && k0 + 32 < n) dv -= fabs(ge[k0 + 31]);
  ds[w][lane] = dv;
  es[w][lane] = lane < 31 ? ge[k0 + lane] : 0.0;
  if (lane == 0) es[w][32] = 0.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) Zs[w][lane][c] = (c == lane) ? 1.0 : 0.0;
  __syncwarp();
  double* d = ds[w];
  double* e = es[w];
  const int nn = DC_LEAF;
  bool fail = false;
  for (int l = 0; l < nn; ++l) {
    int iter = 0;
    for (;;) {
      int mm = l;
      while (mm < nn - 1) {
        if (fabs(e[mm]) <= DBL_EPSILON * (fabs(d[mm]) + fabs(d[mm + 1]))) break;
        ++mm;
      }
      if (mm == l) break;
      if (++iter > 60) { fail = true; break; }
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = hypot(g, 1.0);
      g = d[mm] - d[l] + e[l] / (g + copysign(r, g));
      double s = 1.0, c = 1.0, p = 0.0;
      bool under = false;
      for (int i = mm - 1; i >= l; --i) {
        const double f = s * e[i], b = c * e[i];
        r = hypot(f, g);
        __syncwarp();
        if (lane == 0) e[i + 1] = r;
        if (r == 0.0) {
          if (lane == 0) { d[i + 1] -= p; e[mm] = 0.0; }
          under = true;
          __syncwarp();
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        __syncwarp();
        if (lane == 0) d[i + 1] = g + p;
        g = c * r - b;
        const double zi = Zs[w][lane][i], zi1 = Zs[w][lane][i + 1];
        Zs[w][lane][i + 1] = s * zi + c * zi1;
        Zs[w][lane][i] = c * zi - s * zi1;
        __syncwarp();
      }
      if (under) continue;
      __syncwarp();
      if (lane == 0) { d[l] -= p; e[l] = g; e[mm] = 0.0; }
      __syncwarp();
    }
    if (fail) break;
  }
  if (fail && lane == 0) atomicOr(&info[z * 4 + 1], 4);
  __syncwarp();
  ws.dm[k0 + lane] = d[lane];
  double* q = Q + (size_t)z * qstride + (size_t)(k0 + lane) * ldq + k0;
#pragma unroll
  for (int c = 0; c < 32; ++c) q[c] = Zs[w][lane][c];
}

// ------------------------------------------------------------------------------------------------
// Merge preparation.  grid (n / k, nz), 256 threads, shared memory: d[k], z[k] (doubles), rank/perm [k] (ints).
__global__ void __launch_bounds__(256) dc_prepare_kernel(const double* __restrict__ ee, double* __restrict__ scratch,
                                                         const double* __restrict__ Q, long long qstride, int ldq, int n,
                                                         int k) {
  extern __shared__ __align__(16) double psm[];
  double* sd = psm;
  double* sz = sd + k;
  double* sds = sz + k;
  double* szs = sds + k;
  int* sperm = reinterpret_cast<int*>(szs + k);
  __shared__ double red[40];
  const int m_ = blockIdx.x, z = blockIdx.y, tid = threadIdx.x, h = k / 2;
  const int c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  const double ecut = ee[(size_t)z * n + c0 + h - 1];
  const double sgn = ecut >= 0.0 ? 1.0 : -1.0;
  const double* q = Q + (size_t)z * qstride;
  double ss = 0.0;
  for (int j = tid; j < k; j += 256) {
    sd[j] = ws.dm[c0 + j];
    const double v = j < h ? q[(size_t)(c0 + h - 1) * ldq + c0 + j] : sgn * q[(size_t)(c0 + h) * ldq + c0 + j];
    sz[j] = v;
    ss = fma(v, v, ss);
  }
  ss = block_sum(ss, red);
  const double nrm = sqrt(ss), inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  const double rho = fabs(ecut) * ss;
  __syncthreads();
  // rank sort by counting (ties by index)
  double dmax = 0.0, zmax = 0.0;
  for (int j = tid; j < k; j += 256) {
    const double dj = sd[j];
    int r = 0;
    for (int i = 0; i < k; ++i) {
      const double di = sd[i];
      r += (di < dj) || (di == dj && i < j);
    }
    sperm[r] = j;
    sds[r] = dj;
    szs[r] = sz[j] * inv;
    dmax = fmax(dmax, fabs(dj));
    zmax = fmax(zmax, fabs(sz[j] * inv));
  }
  __syncthreads();
  // max |d|, max |z| (reuse the block reduction through sums of maxima: two passes of warp max)
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = dmax; red[8 + (tid >> 5)] = zmax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { red[0] = fmax(red[0], red[w]); red[8] = fmax(red[8], red[8 + w]); }
    const double tol = 8.0 * DBL_EPSILON * fmax(red[0], red[8]);
    // deflation scan (sequential, like LAPACK dlaed2): negligible weights; close poles by a Givens rotation
    int nk = 0, nd = 0, nr = 0, p = -1;
    for (int j = 0; j < k; ++j) {
      if (rho * fabs(szs[j]) <= tol) {
        ws.dcol[c0 + nd++] = j;
        szs[j] = 0.0;
        continue;
      }
      if (p >= 0) {
        double s_ = szs[p], c_ = szs[j];
        const double tau = hypot(c_, s_);
        const double t = sds[j] - sds[p];
        c_ /= tau;
        s_ = -s_ / tau;
        if (fabs(t * c_ * s_) <= tol) {
          szs[j] = tau;
          szs[p] = 0.0;
          ws.rotp[c0 + nr] = sperm[p]; ws.rotj[c0 + nr] = sperm[j]; ws.rotc[c0 + nr] = c_; ws.rots[c0 + nr] = s_;
          ++nr;
          const double t2 = sds[p] * c_ * c_ + sds[j] * s_ * s_;
          sds[j] = sds[p] * s_ * s_ + sds[j] * c_ * c_;
          sds[p] = t2;
          ws.dcol[c0 + nd++] = p;
          p = j;
          continue;
        }
        ws.keep[c0 + nk++] = p;
      }
      p = j;
    }
    if (p >= 0) ws.keep[c0 + nk++] = p;
    int* meta = ws.meta + (size_t)(c0 / 64) * 4;
    meta[0] = nk; meta[1] = nd; meta[2] = nr; meta[3] = 0;
    ws.rho[(c0 / 64) * 2] = rho;
    red[16] = (double)nk;
  }
  __syncthreads();
  const int nk = (int)red[16];
  const int nd = k - nk;
  for (int j = tid; j < k; j += 256) { ws.perm[c0 + j] = sperm[j]; ws.ds[c0 + j] = sds[j]; ws.zs[c0 + j] = szs[j]; }
  // the secular problem: kept poles, their weights; deflated eigenvalues occupy the first nd columns of U~
  double wsum = 0.0;
  for (int i = tid; i < nk; i += 256) {
    const int s = ws.keep[c0 + i];
    const double zv = szs[s];
    ws.dd[c0 + i] = sds[s];
    ws.zz[c0 + i] = zv;
    ws.ww[c0 + i] = rho * zv * zv;
    wsum += rho * zv * zv;
  }
  for (int i = tid; i < nd; i += 256) ws.lamn[c0 + i] = sds[ws.dcol[c0 + i]];
  wsum = block_sum(wsum, red);
  if (tid == 0) ws.rho[(c0 / 64) * 2 + 1] = wsum;
}

// ------------------------------------------------------------------------------------------------
// One secular root per WARP (the lanes split the m terms of every function evaluation; a thread per root ran 100 x m
// dependent divisions on 2 k threads in all: 19 of the 30 ms of the eigen phase at n = 1024).
// grid (ceil(k / 32), n / k, nz), 256 threads = 8 warps x 4 roots; shared memory dd[k], ww[k].
__global__ void __launch_bounds__(256) dc_secular_kernel(double* __restrict__ scratch, int n, int k, int* __restrict__ info) {
  extern __shared__ __align__(16) double ssm[];
  double* dd = ssm;
  double* ww = ssm + k;
  const int m_ = blockIdx.y, z = blockIdx.z, c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  const int m = ws.meta[(size_t)(c0 / 64) * 4];
  if ((int)blockIdx.x * 32 >= m) return;
  for (int j = threadIdx.x; j < m; j += 256) { dd[j] = ws.dd[c0 + j]; ww[j] = ws.ww[c0 + j]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double rho_total = ws.rho[(c0 / 64) * 2 + 1];
  const int nd = k - m;
  for (int q4 = 0; q4 < 4; ++q4) {
    const int i = blockIdx.x * 32 + warp * 4 + q4;
    if (i >= m) break;                       // (uniform over the warp)
    const bool last = i == m - 1;
    int K;
    double lo, hi;
    if (last) {
      K = i; lo = 0.0; hi = rho_total;
    } else {
      const double gap = dd[i + 1] - dd[i], mid = 0.5 * gap, di = dd[i];
      double fm = 0.0;
      for (int j = lane; j < m; j += 32) fm += ww[j] / ((dd[j] - di) - mid);
      fm = 1.0 + warp_sum(fm);
      if (fm >= 0.0) { K = i; lo = 0.0; hi = mid; }
      else { K = i + 1; lo = -mid; hi = 0.0; }
    }
    const double dK = dd[K];
    // "Middle way" iteration (Li 1994, the scheme of LAPACK dlaed4): the poles to the left of the root (psi) and to the
    // right (phi) are each replaced by ONE pole at the interval end, matched in value and derivative, and the resulting
    // two-pole equation is solved exactly (a quadratic); bracketed, bisection when a step leaves the bracket.
    // ~7 iterations on average instead of ~28 for Newton on the pole-free form (scripts/proto_dc.py).
    double t = 0.5 * (lo + hi);
    bool ok = false;
    for (int it = 0; it < 400; ++it) {
      double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
      for (int j = lane; j < m; j += 32) {
        const double qv = 1.0 / ((dd[j] - dK) - t);
        const double wq = ww[j] * qv;
        if (j <= i) { psi += wq; dpsi = fma(wq, qv, dpsi); }
        else { phi += wq; dphi = fma(wq, qv, dphi); }
      }
      psi = warp_sum(psi); dpsi = warp_sum(dpsi); phi = warp_sum(phi); dphi = warp_sum(dphi);
      const double f = 1.0 + psi + phi;
      if (f == 0.0) { ok = true; break; }
      if (f > 0.0) hi = t; else lo = t;
      const double dl = (dd[i] - dK) - t;                       // delta_i - t  (< 0)
      const double qq = dpsi * dl * dl, pp = psi - dpsi * dl;
      double tn;
      if (last) {
        const double c = 1.0 + pp;
        tn = c != 0.0 ? (dd[i] - dK) + qq / c : 0.5 * (lo + hi);
      } else {
        const double dr = (dd[i + 1] - dK) - t;                 // delta_{i+1} - t  (> 0)
        const double ss = dphi * dr * dr, rr = phi - dphi * dr;
        const double c = 1.0 + pp + rr;
        // c (dl - eta)(dr - eta) + qq (dr - eta) + ss (dl - eta) = 0,  eta = x - t
        const double qa = c, qb = -(c * (dl + dr) + qq + ss), qc = c * dl * dr + qq * dr + ss * dl;
        double eta;
        if (qa == 0.0) {
          eta = qb != 0.0 ? -qc / qb : 0.0;
        } else {
          const double disc = fmax(qb * qb - 4.0 * qa * qc, 0.0), sq = sqrt(disc);
          const double e1 = qb <= 0.0 ? (-qb + sq) / (2.0 * qa) : (-qb - sq) / (2.0 * qa);
          const double e2 = e1 != 0.0 ? qc / (qa * e1) : 0.0;
          eta = (dl < e1 && e1 < dr) ? e1 : e2;
        }
        tn = t + eta;
      }
      if (!(lo < tn && tn < hi)) {
        // bisection; geometric while the bracket still spans orders of magnitude (a root very close to its pole)
        if (lo == 0.0) tn = hi * 0.015625;
        else if (hi == 0.0) tn = lo * 0.015625;
        else if (lo > 0.0 && hi > 4.0 * lo) tn = sqrt(lo * hi);
        else if (hi < 0.0 && lo < 4.0 * hi) tn = -sqrt(lo * hi);
        else tn = 0.5 * (lo + hi);
      }
      if (tn == t || fabs(tn - t) <= 2.0 * DBL_EPSILON * fabs(tn)) { t = tn; ok = true; break; }
      t = tn;
      if (hi - lo <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi))) { ok = true; break; }
    }
    if (lane == 0) {
      if (!ok) atomicOr(&info[z * 4 + 1], 8);
      ws.Ks[c0 + i] = K;
      ws.taus[c0 + i] = t;
      ws.lamn[c0 + nd + i] = dK + t;
    }
  }
}

// Loewner recomputation: rho zhat_j^2 = prod_i (lambda_i - d_j) / prod_{i != j} (d_i - d_j).   thread per j.
__global__ void __launch_bounds__(128) dc_zhat_kernel(double* __restrict__ scratch, int n, int k) {
  extern __shared__ __align__(16) double zsm[];
  double* dd = zsm;
  double* dk = zsm + k;       // d_{K_i}
  double* ta = zsm + 2 * k;   // tau_i
  const int m_ = blockIdx.y, z = blockIdx.z, c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  const int m = ws.meta[(size_t)(c0 / 64) * 4];
  if ((int)blockIdx.x * 128 >= m) return;
  for (int j = threadIdx.x; j < m; j += 128) {
    dd[j] = ws.dd[c0 + j];
    ta[j] = ws.taus[c0 + j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += 128) dk[j] = dd[ws.Ks[c0 + j]];
  __syncthreads();
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= m) return;
  const double dj = dd[j];
  double prod = (dk[j] - dj) + ta[j];
  for (int i = 0; i < m; ++i) {
    if (i == j) continue;
    prod *= ((dk[i] - dj) + ta[i]) / (dd[i] - dj);
  }
  const double rho = ws.rho[(c0 / 64) * 2];
  ws.zhat[c0 + j] = copysign(sqrt(fabs(prod) / rho), ws.zz[c0 + j]);
}

// Column order of the merged block: deflated eigenvalues first, then the roots; at the FINAL level the columns are
// ranked in descending order of the eigenvalue (the order the filters consume them in).   grid (n / k, nz)
__global__ void __launch_bounds__(256) dc_order_kernel(double* __restrict__ scratch, int n, int k, int final_level) {
  const int m_ = blockIdx.x, z = blockIdx.y, c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  for (int c = threadIdx.x; c < k; c += 256) {
    int col = c;
    if (final_level) {
      const double lc = ws.lamn[c0 + c];
      int r = 0;
      for (int i = 0; i < k; ++i) {
        const double li = ws.lamn[c0 + i];
        r += (li > lc) || (li == lc && i < c);
      }
      col = r;
    }
    ws.newcol[c0 + c] = col;
  }
}

// U~^T (stored [new column][original column], k x k per merge).  Deflated column c: unit vector at perm[dcol[c]];
// root i: zhat_j / (d_j - lambda_i), normalised.   grid (k, n / k, nz): one CTA of 128 threads per new column.
__global__ void __launch_bounds__(128) dc_buildu_kernel(double* __restrict__ scratch, double* __restrict__ Ut, long long ustride,
                                                        int n, int k) {
  __shared__ double red[40];
  const int c = blockIdx.x, m_ = blockIdx.y, z = blockIdx.z, c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  const int m = ws.meta[(size_t)(c0 / 64) * 4], nd = k - m;
  double* row = Ut + (size_t)z * ustride + (size_t)m_ * k * k + (size_t)ws.newcol[c0 + c] * k;
  if (c < nd) {
    if (threadIdx.x == 0) row[ws.perm[c0 + ws.dcol[c0 + c]]] = 1.0;
    return;
  }
  const int i = c - nd;
  const double dK = ws.dd[c0 + ws.Ks[c0 + i]], tau = ws.taus[c0 + i];
  double ss = 0.0;
  for (int j = threadIdx.x; j < m; j += 128) {
    const double u = ws.zhat[c0 + j] / ((ws.dd[c0 + j] - dK) - tau);
    ss = fma(u, u, ss);
  }
  ss = block_sum(ss, red);
  const double inv = 1.0 / sqrt(ss);
  for (int j = threadIdx.x; j < m; j += 128) {
    const double u = ws.zhat[c0 + j] / ((ws.dd[c0 + j] - dK) - tau);
    row[ws.perm[c0 + ws.keep[c0 + j]]] = u * inv;
  }
}

// The Givens rotations of the deflation act on columns of Q; Q G U~ = Q (G U~): they are applied to the ROWS of U~
// (= columns of the stored U~^T), last rotation first.   grid (n / k, nz), 256 threads over the new columns.
__global__ void __launch_bounds__(256) dc_rotate_kernel(double* __restrict__ scratch, double* __restrict__ Ut, long long ustride,
                                                        int n, int k) {
  const int m_ = blockIdx.x, z = blockIdx.y, c0 = m_ * k;
  DcWs ws = dc_ws(scratch, n, z);
  const int nr = ws.meta[(size_t)(c0 / 64) * 4 + 2];
  if (nr == 0) return;
  double* U = Ut + (size_t)z * ustride + (size_t)m_ * k * k;
  for (int c = threadIdx.x; c < k; c += 256) {
    double* row = U + (size_t)c * k;          // new column c of U~
    for (int r = nr - 1; r >= 0; --r) {
      const int p = ws.rotp[c0 + r], j = ws.rotj[c0 + r];
      const double cc = ws.rotc[c0 + r], s = ws.rots[c0 + r];
      const double up = row[p], uj = row[j];
      row[p] = cc * up - s * uj;
      row[j] = s * up + cc * uj;
    }
  }
}

__global__ void dc_copy_lam_kernel(double* __restrict__ scratch, double* __restrict__ lam, int n, int k, int V, int final_level) {
  const int z = blockIdx.y;
  DcWs ws = dc_ws(scratch, n, z);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (final_level) {
    const int col = ws.newcol[e];
    if (col < V) lam[(size_t)z * V + col] = ws.lamn[e];
  }
  ws.dm[e] = ws.lamn[e];
}

}  // namespace

bool dc_supported(int n) {
  if (n < 2 * DC_LEAF) return false;
  int m = n / DC_LEAF;
  return n % DC_LEAF == 0 && (m & (m - 1)) == 0;
}

size_t dc_scratch_doubles(int n, int nz) {
  return (size_t)nz * (dc_doubles_per_zone(n) + (dc_ints_per_zone(n) + 1) / 2) + 16;
}

// All n eigenpairs of T (ws.dd, ws.ee): eigenvalues -> ws.lam (descending, first V), eigenvectors -> slot 4 of the
// inverse-iteration workspace (n x Vp row-major, column v = eigenvector of the v-th largest eigenvalue).
// Needs V == n (Vp >= n): scratch = Cm (free after the tridiagonalisation) and slots 0, 1 of the workspace.
int dc_run(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, nz = ws.nz, V = ws.V, Vp = ws.Vp;
  if (!dc_supported(n) || Vp < n) return EINVAL_;
  const size_t need = dc_scratch_doubles(n, nz);
  if (ws.dcw_count < need) {
    if (ws.dcw) cudaFree(ws.dcw);
    ws.dcw = nullptr; ws.dcw_count = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&ws.dcw, need * sizeof(double)));
    APV_CUDA_TRY(cudaMemsetAsync(ws.dcw, 0, need * sizeof(double), st));
    ws.dcw_count = need;
  }
  const long long zs = 6LL * n * Vp;             // zone stride of the inverse-iteration workspace
  double* Qa = ws.Cm;                            // [nz][n][ldn]
  const long long qas = (long long)n * ws.ldn;
  double* Qb = ws.iv;                            // slot 0: [n][Vp]
  double* Ut = ws.iv + (size_t)n * Vp;           // slot 1: merge matrices, n * k doubles per zone
  double* X = ws.iv + 4 * (size_t)n * Vp;        // slot 4: the result
  int nl = 0;
  // leaves -> the diagonal 32 x 32 blocks of Qa (only the diagonal blocks of a level are ever read)
  dc_leaf_kernel<<<dim3(ceil_div(n / DC_LEAF, 4), nz), 128, 0, st>>>(ws.dd, ws.ee, ws.dcw, Qa, qas, ws.ldn, n, ws.info);
  ++nl;
  const double* src = Qa; long long sstride = qas; int sld = ws.ldn;
  bool src_is_a = true;
  for (int k = 2 * DC_LEAF; k <= n; k *= 2) {
    const int nm = n / k;
    const bool fin = k == n;
    const size_t psm = (size_t)4 * k * sizeof(double) + (size_t)k * sizeof(int);
    static PerDevice pd_p; size_t& cp = pd_p.cur();
    if (psm > 48 * 1024 && psm > cp) {
      APV_CUDA_TRY(cudaFuncSetAttribute(dc_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
      cp = psm;
    }
    dc_prepare_kernel<<<dim3(nm, nz), 256, psm, st>>>(ws.ee, ws.dcw, src, sstride, sld, n, k);
    const size_t ssm = (size_t)2 * k * sizeof(double), zsm = (size_t)3 * k * sizeof(double);
    static PerDevice pd_s; size_t& cs = pd_s.cur();
    if (zsm > 48 * 1024 && zsm > cs) {
      APV_CUDA_TRY(cudaFuncSetAttribute(dc_secular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsm));
      APV_CUDA_TRY(cudaFuncSetAttribute(dc_zhat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsm));
      cs = zsm;
    }
    dc_secular_kernel<<<dim3(ceil_div(k, 32), nm, nz), 256, ssm, st>>>(ws.dcw, n, k, ws.info);
    dc_zhat_kernel<<<dim3(ceil_div(k, 128), nm, nz), 128, zsm, st>>>(ws.dcw, n, k);
    dc_order_kernel<<<dim3(nm, nz), 256, 0, st>>>(ws.dcw, n, k, fin ? 1 : 0);
    for (int z = 0; z < nz; ++z)
      APV_CUDA_TRY(cudaMemsetAsync(Ut + (size_t)z * zs, 0, (size_t)n * k * sizeof(double), st));
    dc_buildu_kernel<<<dim3(k, nm, nz), 128, 0, st>>>(ws.dcw, Ut, zs, n, k);
    dc_rotate_kernel<<<dim3(nm, nz), 256, 0, st>>>(ws.dcw, Ut, zs, n, k);
    dc_copy_lam_kernel<<<dim3(ceil_div(n, 256), nz), 256, 0, st>>>(ws.dcw, ws.lam, n, k, V, fin ? 1 : 0);
    nl += 7;
    // Q_new = blockdiag(Q_1, Q_2) U~: per zone one batched GEMM, batch = merges, split = the two halves
    double* dst; long long dstride; int dld;
    if (fin) { dst = X; dstride = zs; dld = Vp; }
    else if (src_is_a) { dst = Qb; dstride = zs; dld = Vp; }
    else { dst = Qa; dstride = qas; dld = ws.ldn; }
    for (int z = 0; z < nz; ++z) {
      GemmArgs g{};
      g.batch = nm; g.split = 2;
      g.A = src + (size_t)z * sstride; g.lda = sld; g.strideA = (long long)k * sld + k; g.splitA = (long long)(k / 2) * sld + k / 2;
      g.B = Ut + (size_t)z * zs; g.ldb = k; g.strideB = (long long)k * k; g.splitB = k / 2; g.transB = 1;
      g.C = dst + (size_t)z * dstride; g.ldc = dld; g.strideC = (long long)k * dld + k; g.splitC = (long long)(k / 2) * dld;
      g.M = k / 2; g.N = fin ? V : k; g.K = k / 2; g.alpha = 1.0; g.beta = 0.0;
      APV_TRY(gemm_f64(g, st));
      ++nl;
    }
    src = dst; sstride = dstride; sld = dld;
    src_is_a = !src_is_a;
  }
  APV_CUDA_TRY(cudaGetLastError());
  if (launches) *launches += nl;
  return OK;
}

}  // namespace apv
