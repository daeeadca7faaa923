// syevd_f64, stage 1: blocked Householder tridiagonalisation  C = Q T Q^T  of the symmetrised
// C = L^-1 R_B L^-T (the reference reaches the same eigen-decomposition through LAPACK dgees,
// Python/apvast.py:30).  Lower variant of LAPACK dsytrd/dlatrd on a fully stored symmetric matrix,
// panel width NBT = 32, both zone problems in the same launches.
//
// Panel storage: Z1[i][0..NBT) = V, Z1[i][NBT..2NBT) = W;  Z2[i][0..NBT) = W, Z2[i][NBT..2NBT) = V, so that
// the trailing update A22 -= V W^T + W V^T is one DMMA GEMM  A22 -= Z1 Z2^T  with K = 2 NBT (issued by
// the host between panel kernels).
//
// ONE persistent cooperative kernel per panel of NBT columns, two barriers per column (per zone: the CTAs of a
// zone synchronise among themselves through a monotone arrival counter, so the zones may drift apart):
//
//   phase P1(j): (jj > 0) raw w of the previous column,  w = tau (y - V (W^T v) - W (V^T v)),  its
//                partial w.v, and the updated column j as a function of the still unknown
//                gamma = tau/2 (w.v):   x = a + 2 gamma v_prev   (a goes to colbuf)
//   --- zone barrier ---
//   phase P2(j): gamma, x and the reflector scalars (beta, tau, s = 1/(alpha-beta)) -- every CTA
//                redundantly, it needs all of x for its GEMV rows anyway --, final W column of the
//                previous reflector, y = A[j+1:, j+1:] v for the CTA's rows (the HBM/L2-bound part:
//                two rows per warp trip, 16 x 16-byte loads in flight per lane, 16 warps per SM -- with 4 in
//                flight the product was limited by per-SM memory-level parallelism, not by HBM),
//                partial (panel)^T v.
//                v = s x + c e_{j+1} (c = 1 - s alpha) is not materialised for the product:
//                A v = s (A x) + c A[:, j+1], so no rescaling pass sits in front of the big read.
//   --- grid barrier ---
// CTAs are split between the zone problems (zone = blockIdx.x % nz).  Everything another CTA wrote
// during the kernel is read with ld.global.cg (L2), never through the non-coherent L1; independent
// loads are issued ahead of the reductions they do not depend on (the phases are latency chains).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace apv {

namespace {

constexpr int NBT = 32;         // panel width (one V and one W column per lane)
constexpr int GMAX = 256;       // upper bound of CTAs per zone
constexpr int TDT = 512;        // threads per CTA
constexpr int TDW = TDT / 32;
constexpr int TD_TS = 0, TD_TRED = 2 * NBT, TD_RED = TD_TRED + 8 * 2 * NBT, TD_SC = TD_RED + 40, TD_EXTRA = TD_SC + 8;

struct TdPanel {
  double* Cm; double* VH; double* Z1; double* Z2; double* tau; double* dd; double* ee;
  double* colbuf; double* ybuf; double* wbuf;
  double* vcur;   // [nz][2][n]        current reflector, double buffered by column parity
  double* pwv;    // [nz][GMAX]        partial w.v
  double* tpart;  // [nz][GMAX][2 NBT] partial (panel col)^T v
  unsigned long long* bar;   // [nz] monotone arrival counters of the per-zone barriers
  int n, ldn, nz;
  long long* dbg;
};

#define TD_TICK(k) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) { long long _t = clock64(); a.dbg[k] += _t - tk; tk = _t; } } while (0)

// Barrier among the G CTAs of ONE zone problem (all CTAs are co-resident: cooperative launch).  The two zones
// synchronise independently, so one zone's latency-bound phases overlap the other zone's HBM-bound product instead
// of both zones idling together.  Monotone 64-bit arrival counter: barrier k completes when it reaches (k+1) G.
__device__ __forceinline__ void zone_sync(unsigned long long* bar, int G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(bar, 1ull);
    const unsigned long long target = (old / (unsigned long long)G + 1ull) * (unsigned long long)G;
    while (*reinterpret_cast<volatile unsigned long long*>(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void chunk_of(int lo, int hi, int G, int g, int& a, int& b) {
  const int rows = hi - lo, per = (rows + G - 1) / G;
  a = lo + g * per;
  b = min(hi, a + per);
}

__device__ __forceinline__ void td_phase1(const TdPanel& a, double* ex, int z, int g, int G, int j, int jj,
                                          bool w_only) {
  const int n = a.n, ldn = a.ldn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ts = ex + TD_TS;
  double* tred = ex + TD_TRED;
  double* red = ex + TD_RED;
  const double* Z1 = a.Z1 + (size_t)z * n * 2 * NBT;
  const double* Cm = a.Cm + (size_t)z * n * ldn;
  double* colbuf = a.colbuf + (size_t)z * n;
  double* wbuf = a.wbuf + (size_t)z * n;
  const double* ybuf = a.ybuf + (size_t)z * n;
  const double* vprev = a.vcur + ((size_t)z * 2 + ((jj + 1) & 1)) * n;   // written in P2 of column j-1
  int r0, r1;
  chunk_of(j, n, G, g, r0, r1);
  if (jj == 0) {
    for (int i = r0 + threadIdx.x; i < r1; i += TDT) colbuf[i] = Cm[(size_t)j * ldn + i];
    return;
  }
  const int nc = jj - 1;                       // finished panel columns
  // loads that do not depend on the reduction: row-j panel entries, y_j, tau
  double Vj = 0.0, Wj = 0.0;
  if (lane < nc) {
    Vj = __ldcg(Z1 + (size_t)j * 2 * NBT + lane);
    Wj = __ldcg(Z1 + (size_t)j * 2 * NBT + NBT + lane);
  }
  const double yj = __ldcg(ybuf + j);
  const double tau_p = __ldcg(a.tau + (size_t)z * n + j - 1);
  {  // reduce the partial panel^T v of column j-1 (columns c' < jj-1 of V and of W)
    const int col = threadIdx.x & 63, qg = threadIdx.x >> 6;       // 8 groups of 64
    double s = 0.0;
    if ((col & (NBT - 1)) < nc) {
      const double* tp = a.tpart + (size_t)z * GMAX * 2 * NBT;
      for (int q = qg; q < G; q += 8) s += __ldcg(tp + (size_t)q * 2 * NBT + col);
    }
    tred[qg * 2 * NBT + col] = s;
    __syncthreads();
    if (threadIdx.x < 2 * NBT) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += tred[q * 2 * NBT + threadIdx.x];
      ts[threadIdx.x] = t;
    }
    __syncthreads();
  }
  double tW = 0.0, tV = 0.0;
  if (lane < nc) { tW = ts[NBT + lane]; tV = ts[lane]; }
  // every CTA recomputes the raw w of row j (V[j][jj-1] = 1)
  const double wj = tau_p * (yj - warp_sum(Vj * tW + Wj * tV));
  double wv = 0.0;
  for (int i0 = r0 + 2 * warp; i0 < r1; i0 += 2 * TDW) {
    // two rows per trip; all loads are issued before the shuffles (same-address loads broadcast)
    const int i1 = i0 + 1;
    const bool ok1 = i1 < r1;
    double V0 = 0.0, W0 = 0.0, V1 = 0.0, W1 = 0.0;
    if (lane < nc) {
      V0 = __ldcg(Z1 + (size_t)i0 * 2 * NBT + lane);
      W0 = __ldcg(Z1 + (size_t)i0 * 2 * NBT + NBT + lane);
      if (ok1) {
        V1 = __ldcg(Z1 + (size_t)i1 * 2 * NBT + lane);
        W1 = __ldcg(Z1 + (size_t)i1 * 2 * NBT + NBT + lane);
      }
    }
    const double y0 = __ldcg(ybuf + i0), vp0 = __ldcg(vprev + i0);
    const double y1 = ok1 ? __ldcg(ybuf + i1) : 0.0, vp1 = ok1 ? __ldcg(vprev + i1) : 0.0;
    double c0 = 0.0, c1 = 0.0;
    if (!w_only) {
      c0 = Cm[(size_t)j * ldn + i0];
      c1 = ok1 ? Cm[(size_t)j * ldn + i1] : 0.0;
    }
    const double cw0 = warp_sum(V0 * tW + W0 * tV), cx0 = warp_sum(V0 * Wj + W0 * Vj);
    const double cw1 = warp_sum(V1 * tW + W1 * tV), cx1 = warp_sum(V1 * Wj + W1 * Vj);
    if (lane == 0) {
      const double w0 = tau_p * (y0 - cw0);
      wbuf[i0] = w0;
      wv += w0 * vp0;
      if (!w_only) colbuf[i0] = c0 - cx0 - vp0 * wj - w0;
      if (ok1) {
        const double w1 = tau_p * (y1 - cw1);
        wbuf[i1] = w1;
        wv += w1 * vp1;
        if (!w_only) colbuf[i1] = c1 - cx1 - vp1 * wj - w1;
      }
    }
  }
  if (lane == 0) red[warp] = wv;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < TDW; ++w) s += red[w];
    a.pwv[(size_t)z * GMAX + g] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ double td_gamma(const TdPanel& a, double* ex, int z, int G, int j, int jj) {
  if (jj == 0) return 0.0;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int q = threadIdx.x; q < G; q += 32) s += __ldcg(a.pwv + (size_t)z * GMAX + q);
    s = warp_sum(s);
    if (threadIdx.x == 0) ex[TD_SC] = 0.5 * __ldcg(a.tau + (size_t)z * a.n + j - 1) * s;
  }
  __syncthreads();
  return ex[TD_SC];
}

__device__ __forceinline__ void td_phase2(const TdPanel& a, double* xs, double* ex, int z, int g, int G, int j,
                                          int jj, long long& tk) {
  const int n = a.n, ldn = a.ldn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tred = ex + TD_TRED;
  double* red = ex + TD_RED;
  double* Z1 = a.Z1 + (size_t)z * n * 2 * NBT;
  double* Z2 = a.Z2 + (size_t)z * n * 2 * NBT;
  const double* Cm = a.Cm + (size_t)z * n * ldn;
  const double* colbuf = a.colbuf + (size_t)z * n;
  const double* wbuf = a.wbuf + (size_t)z * n;
  double* ybuf = a.ybuf + (size_t)z * n;
  const double* vprev = a.vcur + ((size_t)z * 2 + ((jj + 1) & 1)) * n;
  double* vnew = a.vcur + ((size_t)z * 2 + (jj & 1)) * n;
  // the column and the previous reflector are fetched before gamma is known (independent loads)
  constexpr int XPT = 16;                      // elements per thread: n <= XPT * TDT = 8192
  double xa[XPT], xv[XPT];
#pragma unroll
  for (int u = 0; u < XPT; ++u) {
    const int i = j + threadIdx.x + u * TDT;
    xa[u] = (i < n) ? __ldcg(colbuf + i) : 0.0;
    xv[u] = (i < n && jj > 0) ? __ldcg(vprev + i) : 0.0;
  }
  int r0, r1;
  chunk_of(j, n, G, g, r0, r1);
  const double gamma = td_gamma(a, ex, z, G, j, jj);
  double ss = 0.0;
#pragma unroll
  for (int u = 0; u < XPT; ++u) {
    const int i = j + threadIdx.x + u * TDT;
    if (i < n) {
      const double x = fma(2.0 * gamma, xv[u], xa[u]);
      xs[i] = (i == j) ? 0.0 : x;            // xs[j] = 0: the GEMV may start at the even column <= j+1
      if (i == j) ex[TD_SC + 1] = x;         // d_j
      if (i >= j + 2) ss = fma(x, x, ss);
    }
  }
  ss = block_sum(ss, red);                    // (its barriers also publish xs)
  const double dj = ex[TD_SC + 1];
  if (j == n - 1) {
    if (g == 0 && threadIdx.x == 0) a.dd[(size_t)z * n + j] = dj;
    return;
  }
  const double alpha = xs[j + 1];
  double beta, tau, s;
  if (ss == 0.0) {
    beta = alpha; tau = 0.0; s = 0.0;
  } else {
    beta = -copysign(hypot(alpha, sqrt(ss)), alpha);
    tau = (beta - alpha) / beta;
    s = 1.0 / (alpha - beta);
  }
  const double c = 1.0 - s * alpha;            // v = s x + c e_{j+1}
  if (g == 0 && threadIdx.x == 0) {
    a.dd[(size_t)z * n + j] = dj;
    a.ee[(size_t)z * n + j] = beta;
    a.tau[(size_t)z * n + j] = tau;
  }
  double* VHj = a.VH + (size_t)z * n * ldn + (size_t)j * ldn;
  for (int i = r0 + threadIdx.x; i < r1; i += TDT) {
    if (jj > 0) {
      const double wfin = __ldcg(wbuf + i) - gamma * __ldcg(vprev + i);
      Z1[(size_t)i * 2 * NBT + NBT + jj - 1] = wfin;
      Z2[(size_t)i * 2 * NBT + jj - 1] = wfin;
    }
    if (i >= j + 1) {
      const double v = (i == j + 1) ? 1.0 : s * xs[i];
      VHj[i] = v;
      Z1[(size_t)i * 2 * NBT + jj] = v;
      Z2[(size_t)i * 2 * NBT + NBT + jj] = v;
      vnew[i] = v;
    }
  }
  TD_TICK(4);
  // y = s (A x) + c A[:, j+1] for the CTA's rows; columns from the even column c0 <= j+1 (xs[j] = 0, pads are 0)
  const int c0 = (j + 1) & ~1;
  const int rs = max(r0, j + 1);
  const double* rowj1 = Cm + (size_t)(j + 1) * ldn;      // A[i][j+1] = A[j+1][i]
  // two rows per warp trip, 8 independent 16-byte loads per row and lane (16 in flight per lane)
  for (int i = rs + 2 * warp; i < r1; i += 2 * TDW) {
    const bool two = i + 1 < r1;
    const double* row0 = Cm + (size_t)i * ldn;
    const double* row1 = Cm + (size_t)(two ? i + 1 : i) * ldn;
    const double aj0 = rowj1[i], aj1 = two ? rowj1[i + 1] : 0.0;
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    int k = c0 + 2 * lane;
    for (; k + 448 < ldn; k += 512) {
      double2 av[8], bv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) av[u] = __ldg(reinterpret_cast<const double2*>(row0 + k + 64 * u));
#pragma unroll
      for (int u = 0; u < 8; ++u) bv[u] = __ldg(reinterpret_cast<const double2*>(row1 + k + 64 * u));
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        const double2 x0 = *reinterpret_cast<const double2*>(xs + k + 64 * u);
        const double2 x1 = *reinterpret_cast<const double2*>(xs + k + 64 * (u + 1));
        s0 = fma(av[u].x, x0.x, s0); s0 = fma(av[u].y, x0.y, s0);
        s1 = fma(av[u + 1].x, x1.x, s1); s1 = fma(av[u + 1].y, x1.y, s1);
        t0 = fma(bv[u].x, x0.x, t0); t0 = fma(bv[u].y, x0.y, t0);
        t1 = fma(bv[u + 1].x, x1.x, t1); t1 = fma(bv[u + 1].y, x1.y, t1);
      }
    }
    for (; k < ldn; k += 64) {
      const double2 a0 = __ldg(reinterpret_cast<const double2*>(row0 + k));
      const double2 b0 = __ldg(reinterpret_cast<const double2*>(row1 + k));
      const double2 x0 = *reinterpret_cast<const double2*>(xs + k);
      s0 = fma(a0.x, x0.x, s0); s0 = fma(a0.y, x0.y, s0);
      t0 = fma(b0.x, x0.x, t0); t0 = fma(b0.y, x0.y, t0);
    }
    const double ax0 = warp_sum(s0 + s1), ax1 = warp_sum(t0 + t1);
    if (lane == 0) {
      ybuf[i] = fma(s, ax0, c * aj0);
      if (two) ybuf[i + 1] = fma(s, ax1, c * aj1);
    }
  }
  TD_TICK(5);
  __syncthreads();        // the CTA's Z1 writes above are visible to its own threads
  // partial (panel col)^T v over the CTA's rows, columns c' < jj of V and W (W column jj-1 was finalised above)
  {
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    double acc = 0.0;
    if ((col & (NBT - 1)) < jj)
      for (int i = rs + rg; i < r1; i += 8) {
        const double v = (i == j + 1) ? 1.0 : s * xs[i];
        acc = fma(__ldcg(Z1 + (size_t)i * 2 * NBT + col), v, acc);
      }
    tred[rg * 2 * NBT + col] = acc;
    __syncthreads();
    if (threadIdx.x < 2 * NBT) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += tred[q * 2 * NBT + threadIdx.x];
      a.tpart[((size_t)z * GMAX + g) * 2 * NBT + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// end of panel: W[:, pw-1] = w - gamma v for rows >= r
__device__ __forceinline__ void td_phase3(const TdPanel& a, double* ex, int z, int g, int G, int r, int pw) {
  const int n = a.n;
  double* Z1 = a.Z1 + (size_t)z * n * 2 * NBT;
  double* Z2 = a.Z2 + (size_t)z * n * 2 * NBT;
  const double* wbuf = a.wbuf + (size_t)z * n;
  const double* vprev = a.vcur + ((size_t)z * 2 + ((pw + 1) & 1)) * n;
  const double gamma = td_gamma(a, ex, z, G, r, pw);
  int r0, r1;
  chunk_of(r, n, G, g, r0, r1);
  for (int i = r0 + threadIdx.x; i < r1; i += TDT) {
    const double wfin = __ldcg(wbuf + i) - gamma * __ldcg(vprev + i);
    Z1[(size_t)i * 2 * NBT + NBT + pw - 1] = wfin;
    Z2[(size_t)i * 2 * NBT + pw - 1] = wfin;
  }
}

__global__ void __launch_bounds__(TDT, 1) td_panel_kernel(TdPanel a, int k0, int pw) {
  extern __shared__ __align__(16) double td_sm[];
  double* xs = td_sm;                 // [ldn]
  double* ex = td_sm + a.ldn;         // scratch
  const int nz = a.nz, z = blockIdx.x % nz, g = blockIdx.x / nz, G = gridDim.x / nz;
  for (int i = threadIdx.x; i < a.ldn; i += TDT) xs[i] = 0.0;
  __syncthreads();
  long long tk = clock64();
  for (int jj = 0; jj < pw; ++jj) {
    const int j = k0 + jj;
    td_phase1(a, ex, z, g, G, j, jj, false);
    TD_TICK(0);
    zone_sync(a.bar + z, G);
    TD_TICK(1);
    td_phase2(a, xs, ex, z, g, G, j, jj, tk);
    TD_TICK(2);
    zone_sync(a.bar + z, G);
    TD_TICK(3);
  }
  const int r = k0 + pw;
  if (r < a.n) {
    td_phase1(a, ex, z, g, G, r, pw, true);
    zone_sync(a.bar + z, G);
    td_phase3(a, ex, z, g, G, r, pw);
  }
}

}  // namespace

size_t tridiag_scratch_doubles(int n, int nz) {
  return (size_t)nz * GMAX * (2 * NBT + 1) + (size_t)nz * 2 * n + 8;
}

int tridiag_run(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, nz = ws.nz;
  if (n > 16 * TDT) {
    snprintf(g_err, sizeof(g_err), "tridiagonalisation: n = %d exceeds the built limit %d", n, 16 * TDT);
    return EINVAL_;
  }
  int dev = 0, sms = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int G = std::max(1, std::min(std::min(sms / nz, GMAX), ceil_div(n, 16)));
  const size_t smem = (size_t)(ldn + TD_EXTRA) * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (smem > 48 * 1024 && smem > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(td_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  TdPanel tp;
  tp.Cm = ws.Cm; tp.VH = ws.VH; tp.Z1 = ws.Z1; tp.Z2 = ws.Z2; tp.tau = ws.tau; tp.dd = ws.dd; tp.ee = ws.ee;
  tp.colbuf = ws.colbuf; tp.ybuf = ws.ybuf; tp.wbuf = ws.wbuf;
  tp.pwv = ws.tdws;
  tp.tpart = tp.pwv + (size_t)nz * GMAX;
  tp.vcur = tp.tpart + (size_t)nz * GMAX * 2 * NBT;
  tp.bar = reinterpret_cast<unsigned long long*>(tp.vcur + (size_t)nz * 2 * n);   // zeroed at allocation, stays a multiple of G
  tp.n = n; tp.ldn = ldn; tp.nz = nz;
  static long long* dbg = nullptr;
  if (getenv("APV_TD_DEBUG") && !dbg) cudaMalloc((void**)&dbg, 16 * sizeof(long long));
  if (dbg) cudaMemsetAsync(dbg, 0, 16 * sizeof(long long), st);
  tp.dbg = dbg;
  const long long mstride = (long long)n * ldn;
  for (int k0 = 0; k0 < n; k0 += NBT) {
    int pw = std::min(NBT, n - k0), k0v = k0;
    void* args[] = {(void*)&tp, (void*)&k0v, (void*)&pw};
    APV_CUDA_TRY(cudaEventRecord(ws.pev[2 * (k0 / NBT)], st));
    APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)td_panel_kernel, dim3(G * nz), dim3(TDT), args, smem, st));
    APV_CUDA_TRY(cudaEventRecord(ws.pev[2 * (k0 / NBT) + 1], st));
    ++*launches;
    const int r = k0 + pw;
    if (r < n) {         // A22 -= V W^T + W V^T
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
    fprintf(stderr, "td dbg us: P1 %.0f sync1 %.0f P2 %.0f (pre %.0f gemv %.0f) sync2 %.0f\n", hd[0] * us, hd[1] * us,
            hd[2] * us, hd[4] * us, hd[5] * us, hd[3] * us);
  }
  return OK;
}

}  // namespace apv
