// syevd_f64, stage 1: blocked Householder tridiagonalisation  C = Q T Q^T  of the symmetrised
// C = L^-1 R_B L^-T (the reference reaches the same eigen-decomposition through LAPACK dgees,
// Python/apvast.py:30).  Lower variant of LAPACK dsytrd/dlatrd on a fully stored symmetric matrix,
// panel width NBT = 32, both zone problems in the same launches.
//
// ONE persistent cooperative kernel per panel, two grid barriers per column:
//
//   phase A(j)  every CTA gathers the updated column x (both zones) into shared memory and derives the
//               reflector scalars (beta, tau, 1/(alpha-beta)) redundantly.  The symmetric product
//               y' = A x is formed from the LOWER triangle only: the triangle is cut into TB x TB tiles
//               dealt round-robin to the CTAs; a tile yields a row part (y'_I += A_IK x_K) and a column
//               part (y'_K += A_IK^T x_I) that are written as per-tile partial vectors (deterministic,
//               no atomics).  Halving the bytes keeps the trailing matrices of both zones L2-resident
//               for most of the factorisation.  The CTA also accumulates its share of x^T A x and, over
//               its row chunk, the panel products V^T v and W^T v.
//   --- grid barrier ---
//   phase B(j)  one reduction gives V^T v, W^T v and v^T A v, hence
//                   w.v = tau (v^T A v - 2 (V^T v).(W^T v)),   gamma = tau/2 (w.v)
//               without a second reduction; per row of the CTA's chunk: y (assembled from the tile
//               partials), w = tau (y - V W^T v - W V^T v) - gamma v  (final W column), and the next
//               column  x = A[j+1, :] - sum_c (V[:,c] W[j+1,c] + W[:,c] V[j+1,c]).
//   --- grid barrier ---
//
// v = s x + c e_{j+1} (s = 1/(alpha-beta), c = 1 - s alpha) is never materialised for the product:
// A v = s (A x) + c A[:, j+1], so the big read starts before the norm reduction has finished.
// Panel end: A22 -= V W^T + W V^T as one DMMA GEMM (K = 64) issued by the host between panel kernels.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace apv {

namespace {

constexpr int NBT = 32;            // panel width (one V and one W column per lane)
constexpr int TDT = 512;           // threads per CTA
constexpr int TDW = TDT / 32;      // warps per CTA
constexpr int TBMAX = 256;         // largest GEMV tile
constexpr int NTMAX = 33;          // tiles per dimension (n <= 8448 with 256-wide tiles)
constexpr int TILEMAX = NTMAX * (NTMAX + 1) / 2;
constexpr int CTAMAX = 256;        // upper bound of CTAs in the panel kernel
constexpr int NPART = 2 * NBT + 2; // per-CTA partials: V^T v [32], W^T v [32], x^T A x, spare

struct TdArgs {
  double* Cm; double* VH; double* Z1; double* Z2; double* tau; double* dd; double* ee;
  double* xbuf;    // [nz][n]                 updated column j (rows >= j)
  double* pn;      // [nz][CTAMAX]            partial |x[j+2:]|^2
  double* parts;   // [CTAMAX][nz][NPART]
  double* prow;    // [nz][TILEMAX][TBMAX]    tile row parts of A x
  double* pcol;    // [nz][TILEMAX][TBMAX]    tile column parts of A x
  int n, ldn, nz;
  long long* dbg;  // optional clock64 accumulators (APV_TD_DEBUG)
};

#define TD_TICK(k) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) { long long _t = clock64(); a.dbg[k] += _t - *tk; *tk = _t; } } while (0)

// shared-memory scratch layout (doubles) behind xs[nz][ldn] and cb[TDW][TBMAX]
constexpr int EX_TS = 0;                       // reduced V^T v | W^T v           [2 NBT]
constexpr int EX_TRED = EX_TS + 2 * NBT;       // 8-way partials of the above     [8][2 NBT]
constexpr int EX_RED = EX_TRED + 8 * 2 * NBT;  // block_sum scratch               [40]
constexpr int EX_SC = EX_RED + 40;             // per-zone scalars                [2][8]
constexpr int EX_MISC = EX_SC + 16;            // misc                            [16]
constexpr int EX_SIZE = EX_MISC + 16;
enum { SC_SS = 0, SC_DJ, SC_ALPHA, SC_BETA, SC_TAU, SC_S, SC_C };

__device__ __forceinline__ void chunk_of(int lo, int hi, int G, int g, int& a, int& b) {
  const int rows = max(hi - lo, 0), per = (rows + G - 1) / G;
  a = lo + g * per;
  b = min(hi, a + per);
}

__device__ __forceinline__ int tile_size(int m) {    // ~16 tiles per dimension, 64 <= TB <= 256
  int tb = 64;
  while (tb < TBMAX && m > 16 * tb) tb <<= 1;
  return tb;
}

struct Geo {
  int base, tb, nt, ntile;
};
__device__ __forceinline__ Geo geometry(int n, int j) {
  Geo g;
  g.base = (j + 1) & ~1;                 // even, so that 16-byte loads are aligned; x[j] is forced to 0
  g.tb = tile_size(n - j - 1);
  g.nt = (n - g.base + g.tb - 1) / g.tb;
  g.ntile = g.nt * (g.nt + 1) / 2;
  return g;
}

// ---- start of a panel: x = A[k0, k0:n]
__device__ void td_phaseX(const TdArgs& a, double* ex, int k0) {
  const int n = a.n, nz = a.nz;
  const int z = blockIdx.x % nz, g = blockIdx.x / nz, G = gridDim.x / nz;
  if (g >= G) return;
  int r0, r1;
  chunk_of(k0, n, G, g, r0, r1);
  const double* row = a.Cm + (size_t)z * n * a.ldn + (size_t)k0 * a.ldn;
  double ss = 0.0;
  for (int i = r0 + threadIdx.x; i < r1; i += TDT) {
    const double x = row[i];
    a.xbuf[(size_t)z * n + i] = x;
    if (i >= k0 + 2) ss = fma(x, x, ss);
  }
  ss = block_sum(ss, ex + EX_RED);
  if (threadIdx.x == 0) a.pn[(size_t)z * CTAMAX + g] = ss;
}

// ---- phase A
__device__ void td_phaseA(const TdArgs& a, double* xs, double* cb, double* ex, int j, int jj, long long* tk) {
  const int n = a.n, ldn = a.ldn, nz = a.nz;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x, nb = gridDim.x;
  const int zb = b % nz, gb = b / nz, G = nb / nz;
  double* sc = ex + EX_SC;
  // 1. gather x of every zone; x[j] goes to the scalar slot and is zeroed in the vector
  for (int z = 0; z < nz; ++z)
    for (int i = j + threadIdx.x; i < n; i += TDT) {
      const double x = __ldcg(a.xbuf + (size_t)z * n + i);
      if (i == j) { sc[z * 8 + SC_DJ] = x; xs[z * ldn + i] = 0.0; }
      else xs[z * ldn + i] = x;
    }
  if (warp < nz) {
    double s = 0.0;
    for (int q = lane; q < G; q += 32) s += __ldcg(a.pn + (size_t)warp * CTAMAX + q);
    s = warp_sum(s);
    if (lane == 0) sc[warp * 8 + SC_SS] = s;
  }
  __syncthreads();
  if (threadIdx.x < nz) {
    const int z = threadIdx.x;
    const double ss = sc[z * 8 + SC_SS];
    const double alpha = (j + 1 < n) ? xs[z * ldn + j + 1] : 0.0;
    double beta, tau, s;
    if (ss == 0.0) { beta = alpha; tau = 0.0; s = 0.0; }
    else {
      beta = -copysign(hypot(alpha, sqrt(ss)), alpha);
      tau = (beta - alpha) / beta;
      s = 1.0 / (alpha - beta);
    }
    sc[z * 8 + SC_ALPHA] = alpha; sc[z * 8 + SC_BETA] = beta; sc[z * 8 + SC_TAU] = tau;
    sc[z * 8 + SC_S] = s; sc[z * 8 + SC_C] = 1.0 - s * alpha;
    if (b == 0) {
      a.dd[(size_t)z * n + j] = sc[z * 8 + SC_DJ];
      if (j + 1 < n) { a.ee[(size_t)z * n + j] = beta; a.tau[(size_t)z * n + j] = tau; }
    }
  }
  if (j >= n - 1) { __syncthreads(); return; }

  TD_TICK(0);
  // 2. tiles of the lower triangle: row parts, column parts, x^T A x
  const Geo ge = geometry(n, j);
  const int tb = ge.tb, rpw = tb / TDW, q2 = tb / 64;     // rows per warp; double2 per lane and row
  double xax0 = 0.0, xax1 = 0.0;
  for (int tile = b; tile < ge.ntile * nz; tile += nb) {
    const int z = tile / ge.ntile;
    int t = tile - z * ge.ntile, I = 0;
    while (t >= I + 1) { t -= I + 1; ++I; }
    const int K = t;
    const bool offd = I > K;
    const int r0 = ge.base + I * tb, k0 = ge.base + K * tb;
    const double* A = a.Cm + (size_t)z * n * ldn;
    const double* x = xs + z * ldn;
    const int tix = (tile - z * ge.ntile);
    double* prow = a.prow + ((size_t)z * TILEMAX + tix) * TBMAX;
    double colacc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) colacc[u] = 0.0;
    double myx = 0.0;
    for (int rr = 0; rr < rpw; rr += 4) {
      double2 av[4][4];
      double xi[4];
      int ii[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        ii[r] = r0 + warp * rpw + rr + r;
        const bool rowok = (rr + r < rpw) && ii[r] < n;
        xi[r] = rowok ? x[ii[r]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 2 * lane + 64 * u;
          av[r][u] = (rowok && u < q2 && k < ldn && (offd || k <= ii[r]))
                         ? __ldg(reinterpret_cast<const double2*>(A + (size_t)ii[r] * ldn + k))
                         : make_double2(0.0, 0.0);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double racc = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 2 * lane + 64 * u;
          if (u < q2 && k < ldn) {
            const double2 xv = *reinterpret_cast<const double2*>(x + k);
            const bool m0 = offd || k <= ii[r], m1 = offd || k + 1 <= ii[r];
            const bool c0 = offd || k < ii[r], c1 = offd || k + 1 < ii[r];
            racc = fma(m0 ? av[r][u].x : 0.0, xv.x, racc);
            racc = fma(m1 ? av[r][u].y : 0.0, xv.y, racc);
            colacc[2 * u] = fma(c0 ? av[r][u].x : 0.0, xi[r], colacc[2 * u]);
            colacc[2 * u + 1] = fma(c1 ? av[r][u].y : 0.0, xi[r], colacc[2 * u + 1]);
          }
        }
        racc = warp_sum(racc);
        if (lane == 0 && (rr + r < rpw) && ii[r] < n) {
          prow[ii[r] - r0] = racc;
          myx = fma(racc, xi[r], myx);
        }
      }
    }
    // cross-warp reduction of the column parts
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < q2) {
        cb[warp * TBMAX + 2 * lane + 64 * u] = colacc[2 * u];
        cb[warp * TBMAX + 2 * lane + 64 * u + 1] = colacc[2 * u + 1];
      }
    __syncthreads();
    if (threadIdx.x < tb) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < TDW; ++w) s += cb[w * TBMAX + threadIdx.x];
      a.pcol[((size_t)z * TILEMAX + tix) * TBMAX + threadIdx.x] = s;
      const int k = k0 + threadIdx.x;
      if (k < n) myx = fma(s, x[k], myx);
    }
    __syncthreads();
    if (z == 0) xax0 += myx; else xax1 += myx;
  }

  __syncthreads();       // scalars of step 1 visible even to CTAs that own no tile
  TD_TICK(1);
  // 3. panel products over the CTA's row chunk, v stored
  if (gb < G) {
    const double s = sc[zb * 8 + SC_S];
    const double* x = xs + zb * ldn;
    int r0, r1;
    chunk_of(j + 1, n, G, gb, r0, r1);
    double* Z1 = a.Z1 + (size_t)zb * n * 2 * NBT;
    double* Z2 = a.Z2 + (size_t)zb * n * 2 * NBT;
    double* VHj = a.VH + (size_t)zb * n * ldn + (size_t)j * ldn;
    for (int i = r0 + threadIdx.x; i < r1; i += TDT) {
      const double v = (i == j + 1) ? 1.0 : s * x[i];
      VHj[i] = v;
      Z1[(size_t)i * 2 * NBT + jj] = v;
      Z2[(size_t)i * 2 * NBT + NBT + jj] = v;
    }
    double* tred = ex + EX_TRED;
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    double acc = 0.0;
    if ((col & (NBT - 1)) < jj)
      for (int i = r0 + rg; i < r1; i += 8) {
        const double v = (i == j + 1) ? 1.0 : s * x[i];
        acc = fma(__ldcg(Z1 + (size_t)i * 2 * NBT + col), v, acc);
      }
    tred[rg * 2 * NBT + col] = acc;
  }
  // 4. block totals
  double* red = ex + EX_RED;
  const double x0 = block_sum(xax0, red);
  const double x1 = nz > 1 ? block_sum(xax1, red) : 0.0;
  double* P = a.parts + (size_t)b * nz * NPART;
  if (threadIdx.x < 2 * NBT) {
    double t = 0.0;
    if (gb < G) {
#pragma unroll
      for (int q = 0; q < 8; ++q) t += (ex + EX_TRED)[q * 2 * NBT + threadIdx.x];
    }
    P[(size_t)zb * NPART + threadIdx.x] = t;
  }
  if (threadIdx.x == 0) {
    P[2 * NBT] = x0;
    if (nz > 1) P[NPART + 2 * NBT] = x1;
  }
}

// ---- phase B
__device__ void td_phaseB(const TdArgs& a, double* xs, double* ex, int j, int jj, bool last, long long* tk) {
  const int n = a.n, ldn = a.ldn, nz = a.nz;
  if (j >= n - 1) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x, nb = gridDim.x;
  const int zb = b % nz, gb = b / nz, G = nb / nz;
  if (gb >= G) return;
  double* sc = ex + EX_SC + zb * 8;
  double* ts = ex + EX_TS;
  double* tred = ex + EX_TRED;
  double* misc = ex + EX_MISC;
  const double* x = xs + zb * ldn;
  const double* A = a.Cm + (size_t)zb * n * ldn;
  double* Z1 = a.Z1 + (size_t)zb * n * 2 * NBT;
  double* Z2 = a.Z2 + (size_t)zb * n * 2 * NBT;
  const Geo ge = geometry(n, j);
  const double* prow = a.prow + (size_t)zb * TILEMAX * TBMAX;
  const double* pcol = a.pcol + (size_t)zb * TILEMAX * TBMAX;

  // assemble (A x)_i from the tile partial vectors (all lanes of a warp cooperate)
  auto ax_of = [&](int i) -> double {
    const int I = (i - ge.base) / ge.tb, il = (i - ge.base) - I * ge.tb;
    double s = 0.0;
    for (int K = lane; K <= I; K += 32) s += __ldcg(prow + (size_t)(I * (I + 1) / 2 + K) * TBMAX + il);
    for (int I2 = I + lane; I2 < ge.nt; I2 += 32) s += __ldcg(pcol + (size_t)(I2 * (I2 + 1) / 2 + I) * TBMAX + il);
    return warp_sum(s);
  };

  // 1. reductions: V^T v, W^T v over the G CTAs of this zone; x^T A x over all CTAs
  {
    const int col = threadIdx.x & 63, qg = threadIdx.x >> 6;
    double s = 0.0;
    if ((col & (NBT - 1)) < jj)
      for (int q = qg; q < G; q += 8) s += __ldcg(a.parts + ((size_t)(q * nz + zb) * nz + zb) * NPART + col);
    tred[qg * 2 * NBT + col] = s;
    double xa = 0.0;
    for (int q = threadIdx.x; q < nb; q += TDT) xa += __ldcg(a.parts + ((size_t)q * nz + zb) * NPART + 2 * NBT);
    xa = block_sum(xa, ex + EX_RED);          // (barriers inside also publish tred)
    if (threadIdx.x < 2 * NBT) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += tred[q * 2 * NBT + threadIdx.x];
      ts[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) misc[0] = xa;
    __syncthreads();
  }
  TD_TICK(4);
  const double tau = sc[SC_TAU], s = sc[SC_S], c = sc[SC_C];
  const int j1 = j + 1;
  // 2. scalars every thread needs: gamma, and row j+1 of the final panel
  double tW = 0.0, tV = 0.0, Vj1 = 0.0, Wj1 = 0.0;
  if (lane < jj) {
    tV = ts[lane];
    tW = ts[NBT + lane];
    Vj1 = __ldcg(Z1 + (size_t)j1 * 2 * NBT + lane);
    Wj1 = __ldcg(Z1 + (size_t)j1 * 2 * NBT + NBT + lane);
  }
  const double ajj = A[(size_t)j1 * ldn + j1];
  const double axj1 = ax_of(j1);
  const double vAv = s * s * misc[0] + 2.0 * s * c * axj1 + c * c * ajj;
  const double tvtw = warp_sum(tV * tW);
  const double gamma = 0.5 * tau * (tau * (vAv - 2.0 * tvtw));
  const double yj1 = s * axj1 + c * ajj;
  const double wj1 = tau * (yj1 - warp_sum(Vj1 * tW + Wj1 * tV)) - gamma;     // final W[j+1][jj] (v_{j+1} = 1)

  TD_TICK(5);
  // 3. rows of the chunk
  int r0, r1;
  chunk_of(j1, n, G, gb, r0, r1);
  const double* rowj1 = A + (size_t)j1 * ldn;
  double ssn = 0.0;
  for (int i = r0 + warp; i < r1; i += TDW) {
    double Vi = 0.0, Wi = 0.0;
    if (lane < jj) {
      Vi = __ldcg(Z1 + (size_t)i * 2 * NBT + lane);
      Wi = __ldcg(Z1 + (size_t)i * 2 * NBT + NBT + lane);
    }
    const double aji = rowj1[i];                       // A[j+1][i] = A[i][j+1]
    const double axi = ax_of(i);
    const double cw = warp_sum(Vi * tW + Wi * tV);
    const double cx = warp_sum(Vi * Wj1 + Wi * Vj1);
    if (lane == 0) {
      const double vi = (i == j1) ? 1.0 : s * x[i];
      const double yi = s * axi + c * aji;
      const double wf = tau * (yi - cw) - gamma * vi;  // final W[i][jj]
      Z1[(size_t)i * 2 * NBT + NBT + jj] = wf;
      Z2[(size_t)i * 2 * NBT + jj] = wf;
      if (!last) {
        const double xn = aji - cx - vi * wj1 - wf;    // column j+1 after the rank-2 updates (V[j+1][jj] = 1)
        a.xbuf[(size_t)zb * n + i] = xn;
        if (i >= j1 + 2) ssn = fma(xn, xn, ssn);
      }
    }
  }
  if (!last) {
    ssn = block_sum(ssn, ex + EX_RED);
    if (threadIdx.x == 0) a.pn[(size_t)zb * CTAMAX + gb] = ssn;
  }
}

__global__ void __launch_bounds__(TDT, 1) td_panel_kernel(TdArgs a, int k0, int pw) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double td_sm[];
  double* xs = td_sm;                                  // [nz][ldn]
  double* cb = td_sm + (size_t)a.nz * a.ldn;           // [TDW][TBMAX]
  double* ex = cb + TDW * TBMAX;
  for (int i = threadIdx.x; i < a.nz * a.ldn; i += TDT) xs[i] = 0.0;
  __syncthreads();
  long long tkv = clock64();
  long long* tk = &tkv;
  td_phaseX(a, ex, k0);
  grid.sync();
  TD_TICK(8);
  for (int jj = 0; jj < pw; ++jj) {
    const int j = k0 + jj;
    td_phaseA(a, xs, cb, ex, j, jj, tk);
    TD_TICK(2);
    grid.sync();
    TD_TICK(3);
    const bool last = jj == pw - 1;
    td_phaseB(a, xs, ex, j, jj, last, tk);
    TD_TICK(6);
    if (!last) grid.sync();
    TD_TICK(7);
  }
}

}  // namespace

size_t tridiag_scratch_doubles(int n, int nz) {
  (void)n;
  return (size_t)nz * CTAMAX + (size_t)CTAMAX * nz * NPART + 2 * (size_t)nz * TILEMAX * TBMAX;
}

int tridiag_run(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, nz = ws.nz;
  if (ceil_div(n, TBMAX) > NTMAX - 1) {
    snprintf(g_err, sizeof(g_err), "tridiagonalisation: n = %d exceeds the built tile table (n <= %d)", n, (NTMAX - 1) * TBMAX);
    return EINVAL_;
  }
  int dev = 0, sms = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int ctas = std::min(std::min(sms, CTAMAX), std::max(1, ceil_div(n, 8)) * nz);
  ctas = std::max(nz, ctas / nz * nz);
  const size_t smem = ((size_t)nz * ldn + TDW * TBMAX + EX_SIZE) * sizeof(double);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(td_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  TdArgs ta;
  ta.Cm = ws.Cm; ta.VH = ws.VH; ta.Z1 = ws.Z1; ta.Z2 = ws.Z2; ta.tau = ws.tau; ta.dd = ws.dd; ta.ee = ws.ee;
  ta.xbuf = ws.colbuf;
  ta.pn = ws.tdws;
  ta.parts = ta.pn + (size_t)nz * CTAMAX;
  ta.prow = ta.parts + (size_t)CTAMAX * nz * NPART;
  ta.pcol = ta.prow + (size_t)nz * TILEMAX * TBMAX;
  ta.n = n; ta.ldn = ldn; ta.nz = nz;
  static long long* dbg = nullptr;
  if (getenv("APV_TD_DEBUG") && !dbg) { cudaMalloc((void**)&dbg, 16 * sizeof(long long)); }
  if (dbg) cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
  ta.dbg = dbg;
  const long long mstride = (long long)n * ldn;
  for (int k0 = 0; k0 < n; k0 += NBT) {
    int pw = std::min(NBT, n - k0), k0v = k0;
    void* args[] = {(void*)&ta, (void*)&k0v, (void*)&pw};
    APV_CUDA_TRY(cudaEventRecord(ws.pev[2 * (k0 / NBT)], st));
    APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)td_panel_kernel, dim3(ctas), dim3(TDT), args, smem, st));
    APV_CUDA_TRY(cudaEventRecord(ws.pev[2 * (k0 / NBT) + 1], st));
    ++*launches;
    const int r = k0 + pw;
    if (r < n) {         // A22 -= V W^T + W V^T  (Z1 = [V | W], Z2 = [W | V])
      GemmArgs u{};
      u.batch = nz;
      u.A = ws.Z1 + (size_t)r * 2 * NBT; u.lda = 2 * NBT; u.strideA = (long long)n * 2 * NBT;
      u.B = ws.Z2 + (size_t)r * 2 * NBT; u.ldb = 2 * NBT; u.strideB = (long long)n * 2 * NBT;
      u.C = ws.Cm + (size_t)r * ldn + r; u.ldc = ldn; u.strideC = mstride;
      u.M = n - r; u.N = n - r; u.K = 2 * NBT; u.transB = 1; u.alpha = -1.0; u.beta = 1.0;
      APV_TRY(gemm_f64(u, st));
      ++*launches;
    }
  }
  if (dbg) {
    long long hd[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hd, dbg, sizeof(hd), cudaMemcpyDeviceToHost);
    const double us = 1.0 / 1965.0;
    fprintf(stderr, "td dbg us: A.gather %.0f A.tiles %.0f A.rest %.0f sync1 %.0f | B.reduce %.0f B.scalars %.0f B.rows %.0f sync2 %.0f | X %.0f\n",
            hd[0] * us, hd[1] * us, hd[2] * us, hd[3] * us, hd[4] * us, hd[5] * us, hd[6] * us, hd[7] * us, hd[8] * us);
  }
  return OK;
}

}  // namespace apv
