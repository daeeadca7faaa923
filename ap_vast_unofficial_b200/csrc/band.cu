// syevd_f64, two-stage tridiagonalisation (eig_mode 3):  C = Q1 Cb Q1^T = Q1 Q2 T Q2^T Q1^T.
// The reference reaches the eigen-decomposition of C = L^-1 R_B L^-T through LAPACK dgees (Python/apvast.py:30);
// the one-stage reduction of tridiag.cu has to stream the trailing matrix once per column (BLAS-2, HBM-bound).
// Here the O(n^3) work is all BLAS-3 on the FP64 tensor cores and the BLAS-2 part runs on a band that stays in L2:
//
//   stage 1 (dense -> band, half bandwidth NB2 = 32), per panel of NB2 columns
//     sb_panel_qr_kernel   Householder QR of the sub-diagonal panel  C[r:, j0:j0+NB2]  (r = j0 + NB2) spread over one
//                          thread-block cluster per zone, the rows in registers: per column one exchange through
//                          distributed shared memory (st.async into the receiver's transaction barrier, no cluster
//                          barrier); a Gram row of the column gives norm, reflector, the row v^T P and the column of
//                          V^T V needed by the compact-WY factor T in the same reduction
//     gemm (split-K)       Y = C22 V                                       (DMMA, skinny tile)
//     sb_w1/w2_kernel      X = Y T,  W = X - 1/2 V T^T (V^T X),  panels Z1 = [V | W], Z2 = [W | V]
//     gemm                 C22 -= Z1 Z2^T   (= V W^T + W V^T)              (DMMA)
//     look-ahead: the next panel's columns are updated first (skinny GEMM) and its QR runs on a high-priority side
//     stream beside the bulk of the rank-2b update.
//   stage 2 (band -> tridiagonal), sb2st_chase_kernel: bulge chasing.  The critical path is 2 n dependent steps, so
//     the step latency is everything: a CTA owns a group of 3 consecutive sweeps running as a wavefront on a sliding
//     window of band rows in shared memory (two warps per sweep: off-diagonal block / diagonal block, the 32 x 32
//     blocks in registers), a storer and a fetcher warp stream the window and hand the finished rows to the next
//     group through release / acquire row bounds in global memory.  The band (n x 64 doubles per zone) lives in L2.
//   back-transformation of the V wanted eigenvectors: sb2st_apply_q2_kernel (stage-2 reflectors, one CTA per
//     vector, vector in shared memory), then sb_apply_q1_kernel (stage-1 block reflectors with the panel T factors:
//     every CTA keeps a slab of rows of all the vectors in shared memory, the reflector panels are streamed once).
// APV_TS_DEBUG=1 prints per-kernel-class times (synchronising, no look-ahead) and in-kernel clock totals.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace apv {

namespace {

constexpr int NB2 = 32;          // half bandwidth of the intermediate band matrix == warp size (one lane per row)
constexpr int LDB = 2 * NB2;     // band storage (row-major): AB[i][e] = A[i][i - e], e = 0 .. 2 NB2 - 1 (band + bulge)
constexpr int QRT = 1024;        // threads of the panel QR CTAs (32 warps, lane = panel column)
constexpr int QRW = QRT / 32;
constexpr int QR_MAXCS = 16;     // largest cluster
constexpr int PP = NB2 + 1;      // shared-memory pitch of the panel slab (odd: rows and columns conflict-free)
constexpr int PROG_DONE = 0x7fffffff;

// ------------------------------------------------------------------------------------------------
// Stage 1, panel factorisation.  grid (CS, nz), cluster (CS, 1, 1).
struct SbPanel {
  double* Cm; double* VH; double* VP; double* tau; double* Tp;
  int n, ldn, j0, rows_per;
  long long* dbg;      // APV_TS_DEBUG: clock64 totals per phase (rank 0 / warp 0 of zone 0)
};


// ---- distributed-shared-memory hand-over of the panel QR: remote stores that complete a transaction count on the
// RECEIVER's mbarrier (st.async ... mbarrier::complete_tx), so a rank waits on its own barrier for exactly the bytes of
// one column instead of the whole cluster meeting in a barrier.cluster per column
__device__ __forceinline__ unsigned cl_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cl_mapa(const void* p, int rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cl_smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_async(unsigned remote_addr, double v, unsigned remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cl_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cl_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CL_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni CL_WAIT_DONE;\n"
      "bra.uni CL_WAIT_LOOP;\n"
      "CL_WAIT_DONE:\n"
      "}\n" ::"r"(cl_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

#define SB_TICK(k) do { if (dbgp) { const long long _t = clock64(); if (threadIdx.x == 0) dbgp[k] += _t - tk; tk = _t; } } while (0)

// The slab of panel rows lives in REGISTERS: lane = panel column, warp w owns the local rows w, w + 32, .. (RPW of them,
// x[t] = row w + 32 t).  A column step then needs two shuffles per row (the row's entry in column j, and after the
// update its entry in column j + 1 for the next Gram row) and no shared-memory traffic at all; the shared-memory
// version read and wrote every row once per column (7 shared-memory / shuffle wavefronts per row instead of 4, and
// the load -> update -> store -> shuffle chain exposed to their latencies).  The slab goes to shared memory once, at
// the end, to write V out transposed with coalesced stores.
template <int RPW>
__global__ void __launch_bounds__(QRT, 1) sb_panel_qr_kernel(SbPanel a) {
  extern __shared__ __align__(16) double slab[];          // [rows_per][PP] (output staging only)
  __shared__ double xpart[2][QR_MAXCS][NB2];              // per-rank partial Gram rows (written by every rank)
  __shared__ double xrow[2][NB2];                         // the diagonal row of the current column (from its owner)
  __shared__ double red[QRW][NB2];
  __shared__ double Gs[NB2][PP];                          // V^T V above the diagonal (rank 0)
  __shared__ double taus[NB2];
  __shared__ double Tsm[NB2][PP];
  __shared__ double tqs[NB2], scal[2];                    // tau * (v^T P) per column, {1 / (alpha - beta), beta}
  __shared__ __align__(8) unsigned long long colbar[2];   // per parity: the bytes of one column (partials of every rank + its diagonal row)
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int z = blockIdx.y, n = a.n, ldn = a.ldn, j0 = a.j0, r = j0 + NB2, npn = n - r;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Cm = a.Cm + (size_t)z * n * ldn;
  long long* dbgp = (a.dbg && z == 0 && rank == 0) ? a.dbg : nullptr;
  long long tk = clock64();
  const int s0 = min(npn, rank * a.rows_per), s1 = min(npn, s0 + a.rows_per), nr = s1 - s0;   // slab rows [s0, s1)
  if (threadIdx.x == 0) {
    cl_mbar_init(&colbar[0], 1);
    cl_mbar_init(&colbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double x[RPW];
#pragma unroll
  for (int t = 0; t < RPW; ++t) {
    const int i = warp + QRW * t;
    x[t] = (i < nr) ? Cm[(size_t)(r + s0 + i) * ldn + j0 + lane] : 0.0;      // (rows beyond the slab hold zeros)
  }
  cluster.sync();             // every CTA of the cluster runs (and has initialised its barriers) before it is written remotely
  // partial Gram row of column 0 over the rows strictly below its diagonal
  double acc;
  {
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < RPW; ++t) {
      const double x0 = __shfl_sync(0xffffffffu, x[t], 0);
      if (s0 + warp + QRW * t > 0) a4[t & 3] = fma(x[t], x0, a4[t & 3]);
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  SB_TICK(0);
  const bool t0_special = s0 + warp <= NB2;       // (warp-uniform)
  for (int j = 0; j < NB2; ++j) {
    const int par = j & 1;
    red[warp][lane] = acc;
    __syncthreads();
    SB_TICK(1);
    // One column = one transaction phase of colbar[par]: every rank's partial Gram row (CS x 256 bytes) and, if the column
    // has a diagonal row, that row from its owner (256 bytes).  The buffers of parity `par` are rewritten for column
    // j + 2, which a rank sends only after it has received every rank's data of column j + 1 -- sent after those ranks
    // had read column j: no barrier is needed to protect them.
    if (warp == 0 && lane == 0)
      cl_mbar_expect_tx(&colbar[par], (unsigned)((CS + (j < npn ? 1 : 0)) * NB2 * sizeof(double)));
    if (warp < CS) {           // warp w sums the CTA's partial Gram row (redundantly, in parallel) and delivers it to rank w
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
      for (int w = 0; w < QRW; w += 4) {
        t0 += red[w][lane]; t1 += red[w + 1][lane]; t2 += red[w + 2][lane]; t3 += red[w + 3][lane];
      }
      const double t = (t0 + t1) + (t2 + t3);
      cl_st_async(cl_mapa(&xpart[par][rank][lane], warp), t, cl_mapa(&colbar[par], warp));
    }
    // the diagonal row j is the local row j - s0 < 32 of its owner: warp j - s0, register x[0]; that warp hands it to
    // every rank
    if (j >= s0 && j < s1 && warp == j - s0)
      for (int k = 0; k < CS; ++k) cl_st_async(cl_mapa(&xrow[par][lane], k), x[0], cl_mapa(&colbar[par], k));
    SB_TICK(2);
    if (warp == 0) cl_mbar_wait(&colbar[par], (unsigned)((j >> 1) & 1));
    SB_TICK(3);
    if (warp == 0) {           // reflector scalars once per CTA (32 warps doing this redundantly fill the FP64 pipe)
      double g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int k = 0; k < QR_MAXCS; k += 2) {          // (unrolled: the loads of all ranks' partials are in flight together)
        if (k < CS) g0 += xpart[par][k][lane];
        if (k + 1 < CS) g1 += xpart[par][k + 1][lane];
      }
      const double g = g0 + g1;
      const double rowj = (j < npn) ? xrow[par][lane] : 0.0;      // (no diagonal row: the column is empty)
      const double alpha = __shfl_sync(0xffffffffu, rowj, j), sigma = __shfl_sync(0xffffffffu, g, j);
      double beta = alpha, tau = 0.0, scale = 0.0;
      if (sigma > 0.0 && j < npn - 1) {      // (as warp_house below: one rsqrt and one reciprocal)
        const double nrm2 = fma(alpha, alpha, sigma), rs = rsqrt(nrm2);
        beta = -copysign(nrm2 * rs, alpha);
        tau = fma(fabs(alpha), rs, 1.0);
        scale = 1.0 / (alpha - beta);
      }
      const double q = fma(scale, g, rowj);     // lane > j: (v^T P)[lane];  lane < j: (V^T V)[lane][j]
      // per-lane multiplier of the row update x <- x - v_r m: tau (v^T P) for the trailing columns, zero elsewhere
      tqs[lane] = lane > j ? tau * q : 0.0;
      if (lane == 0) { scal[0] = scale; scal[1] = beta; }
      if (rank == 0) {
        if (lane < j) Gs[lane][j] = q;
        if (lane == j) taus[j] = tau;
      }
    }
    __syncthreads();
    const double scale = scal[0], beta = scal[1], m1 = tqs[lane];
    const bool isj = lane == j;
    // rows below the diagonal: x <- x - v_r m, v_r = x_j scale; column j keeps the reflector entry v_r itself, the lanes
    // < j have m = 0.  Written per lane as x s - v_r m' with (s, m') = (scale, 0) in lane j and (1, m) elsewhere: no
    // selects (the kernel is issue-bound: ncu, 57 % issue-active, 35 instructions per row in the branchy form).
    const double sl = isj ? scale : 1.0, ml = isj ? 0.0 : m1;
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    // Only the first row of a warp can still be a finished row, the diagonal row or the next diagonal row (the rows
    // w + 32 t, t >= 1, lie below row 32 of the panel): everything else takes the branch-free form.
    if (t0_special) {
      const int gr = s0 + warp;                       // global row of the panel (warp-uniform)
      if (gr > j) {
        const double vr = __shfl_sync(0xffffffffu, x[0], j) * scale;
        x[0] = fma(-vr, ml, x[0] * sl);
        if (gr > j + 1) {                             // (the next diagonal row takes the update but is not part of the next Gram row)
          const double xb = __shfl_sync(0xffffffffu, x[0], (j + 1) & 31);
          a4[0] = fma(x[0], xb, a4[0]);
        }
      } else if (gr == j) {                           // the diagonal row: R (v_j = 1 on it)
        if (lane > j) x[0] -= m1;
        else if (isj) x[0] = beta;
      }
    }
#pragma unroll
    for (int t = 0; t < RPW; ++t) {
      if (t == 0 && t0_special) continue;
      const double vr = __shfl_sync(0xffffffffu, x[t], j) * scale;
      x[t] = fma(-vr, ml, x[t] * sl);
      const double xb = __shfl_sync(0xffffffffu, x[t], (j + 1) & 31);
      a4[t & 3] = fma(x[t], xb, a4[t & 3]);
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    SB_TICK(5);
  }
  // the slab to shared memory for the outputs
#pragma unroll
  for (int t = 0; t < RPW; ++t) {
    const int i = warp + QRW * t;
    if (i < nr) slab[i * PP + lane] = x[t];
  }
  cluster.sync();              // (also the CTA barrier for the slab) nobody leaves while a remote store into its shared memory could still be under way
  // outputs: R (and zeros) into the panel of C, explicit V (row-major panel and as rows of VH)
  double* VP = a.VP + (size_t)z * n * NB2;
  for (int i = warp; i < nr; i += QRW) {
    const int gr = s0 + i;
    const double xv = slab[i * PP + lane];
    Cm[(size_t)(r + gr) * ldn + j0 + lane] = (gr <= lane) ? xv : 0.0;
    VP[(size_t)(r + gr) * NB2 + lane] = (gr > lane) ? xv : (gr == lane ? 1.0 : 0.0);
  }
  double* VH = a.VH + (size_t)z * n * ldn;
  for (int c = warp; c < NB2; c += QRW)
    for (int i = lane; i < nr; i += 32) {
      const int gr = s0 + i;
      VH[(size_t)(j0 + c) * ldn + r + gr] = (gr > c) ? slab[i * PP + c] : (gr == c ? 1.0 : 0.0);
    }
  if (rank == 0) {
    if (warp == 0) {
      a.tau[(size_t)z * n + j0 + lane] = taus[lane];
      // T = (diag(1 / tau) + striu(V^T V))^-1 (the dlarft factor): lane = column c, back substitution
      //   t[c] = tau_c,  t[r] = -tau_r sum_{k = r+1 .. c} G[r][k] t[k]  (r < c),  zero below the diagonal
      double* Tp = a.Tp + (size_t)z * NB2 * NB2;
      for (int rr = NB2 - 1; rr >= 0; --rr) {
        double a0 = 0.0, a1 = 0.0;
        int k = rr + 1;
        for (; k + 1 < NB2; k += 2) {
          a0 = fma(Gs[rr][k], Tsm[k][lane], a0);                  // (Tsm[k][lane] = 0 for k > lane)
          a1 = fma(Gs[rr][k + 1], Tsm[k + 1][lane], a1);
        }
        if (k < NB2) a0 = fma(Gs[rr][k], Tsm[k][lane], a0);
        const double tt = (rr == lane) ? taus[rr] : (rr < lane ? -taus[rr] * (a0 + a1) : 0.0);
        Tsm[rr][lane] = tt;                                        // (column `lane` is private to this lane)
        Tp[rr * NB2 + lane] = tt;
      }
    }
  }
  SB_TICK(6);
}

// ------------------------------------------------------------------------------------------------
// X = (sum_s Ypart[s]) T  and the partial  V^T X  of a block of 64 rows.   grid (ceil(npn/64), nz), 256 threads.
struct SbW {
  const double* Ypart; const double* VP; const double* Tp;
  double* X; double* Spart; double* Z1; double* Z2;
  int n, r, npn, nsplit, nblk;
  double* Cm; int ldn; int next_cols;     // look-ahead: also update the next panel's NB2 columns of C22 (sb_w2_kernel)
};

__global__ void __launch_bounds__(256) sb_w1_kernel(SbW a) {
  constexpr int WP = NB2 + 4;      // even pitch: 16-byte shared loads of row segments
  __shared__ __align__(16) double Ys[64][WP], Ts[NB2][WP];
  __shared__ double Vs[64][PP];
  const int z = blockIdx.y, blk = blockIdx.x, n = a.n;
  const int row = threadIdx.x >> 2, cq = (threadIdx.x & 3) * 8;
  const int gr = a.r + blk * 64 + row;                       // global row
  const bool ok = gr < n;
  for (int i = threadIdx.x; i < NB2 * NB2; i += 256) Ts[i / NB2][i % NB2] = a.Tp[(size_t)z * NB2 * NB2 + i];
  double y[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) y[c] = 0.0;
  if (ok) {
    const double2* yp = reinterpret_cast<const double2*>(a.Ypart + ((size_t)z * a.nsplit * n + gr) * NB2 + cq);
    const size_t sstride = (size_t)n * NB2 / 2;
#pragma unroll 4
    for (int s = 0; s < a.nsplit; ++s) {       // (all the slices' loads in flight together)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 t = yp[s * sstride + c];
        y[2 * c] += t.x; y[2 * c + 1] += t.y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    Ys[row][cq + c] = y[c];
    Vs[row][cq + c] = ok ? a.VP[((size_t)z * n + gr) * NB2 + cq + c] : 0.0;
  }
  __syncthreads();
  double x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = 0.0;
#pragma unroll 8
  for (int k = 0; k < NB2; ++k) {
    const double yk = Ys[row][k];
    const double2* tr = reinterpret_cast<const double2*>(&Ts[k][cq]);      // T is upper triangular (zeros stored)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double2 t = tr[c];
      x[2 * c] = fma(yk, t.x, x[2 * c]);
      x[2 * c + 1] = fma(yk, t.y, x[2 * c + 1]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    Ys[row][cq + c] = x[c];
    if (ok) a.X[((size_t)z * n + gr) * NB2 + cq + c] = x[c];
  }
  __syncthreads();
  // S_partial[p][c] = sum_rows V[row][p] X[row][c]:  thread -> p = tid / 8, c = (tid % 8) * 4 .. + 3
  const int p = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 4;
  double s4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
  for (int i = 0; i < 64; ++i) {
    const double v = Vs[i][p];
    const double2 xa = *reinterpret_cast<const double2*>(&Ys[i][c0]);
    const double2 xb = *reinterpret_cast<const double2*>(&Ys[i][c0 + 2]);
    s4[0] = fma(v, xa.x, s4[0]); s4[1] = fma(v, xa.y, s4[1]);
    s4[2] = fma(v, xb.x, s4[2]); s4[3] = fma(v, xb.y, s4[3]);
  }
  double* sp = a.Spart + ((size_t)(z * a.nblk + blk) * NB2 + p) * NB2 + c0;
#pragma unroll
  for (int c = 0; c < 4; ++c) sp[c] = s4[c];
}

// W = X - V (1/2 T^T S),  S = sum of the partials;  Z1 = [V | W], Z2 = [W | V].   1024 threads (one entry of S each).
// With next_cols the kernel also applies the rank-2b update to the NB2 leading columns of C22 (the next panel and
// its diagonal block: C22[:, 0:NB2] -= V W0^T + W V0^T with the first NB2 rows W0, V0 of W and V, which every CTA
// recomputes), so that the next panel factorisation can start without waiting for a separate GEMM.
__global__ void __launch_bounds__(1024) sb_w2_kernel(SbW a) {
  extern __shared__ double w2sm[];
  double (*Ss)[PP] = reinterpret_cast<double (*)[PP]>(w2sm);                       // S, later W0
  double (*Ts)[PP] = reinterpret_cast<double (*)[PP]>(w2sm + NB2 * PP);            // T, later V0
  double (*Ms)[PP] = reinterpret_cast<double (*)[PP]>(w2sm + 2 * NB2 * PP);
  double (*Vs)[PP] = reinterpret_cast<double (*)[PP]>(w2sm + 3 * NB2 * PP);        // [64][PP]
  double (*Ws)[PP] = reinterpret_cast<double (*)[PP]>(w2sm + (3 * NB2 + 64) * PP); // [64][PP]
  double (*V0s)[PP] = Ts;
  double (*W0s)[PP] = Ss;
  const int z = blockIdx.y, blk = blockIdx.x, n = a.n, t = threadIdx.x;
  {
    const double* sp = a.Spart + (size_t)z * a.nblk * NB2 * NB2 + t;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = 0;
    for (; b + 15 < a.nblk; b += 16) {          // 16 partials in flight (the loop is a chain of L2 round trips otherwise)
      double t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = sp[(size_t)(b + u) * NB2 * NB2];
#pragma unroll
      for (int u = 0; u < 16; u += 4) { s0 += t[u]; s1 += t[u + 1]; s2 += t[u + 2]; s3 += t[u + 3]; }
    }
    for (; b + 3 < a.nblk; b += 4) {
      s0 += sp[(size_t)b * NB2 * NB2];
      s1 += sp[(size_t)(b + 1) * NB2 * NB2];
      s2 += sp[(size_t)(b + 2) * NB2 * NB2];
      s3 += sp[(size_t)(b + 3) * NB2 * NB2];
    }
    for (; b < a.nblk; ++b) s0 += sp[(size_t)b * NB2 * NB2];
    Ss[t / NB2][t % NB2] = (s0 + s1) + (s2 + s3);
    Ts[t / NB2][t % NB2] = a.Tp[(size_t)z * NB2 * NB2 + t];
  }
  const int row = t >> 4, cq = (t & 15) * 2;          // 64 rows x 16 column pairs
  const int gr = a.r + blk * 64 + row;
  const bool ok = gr < n;
  double x0 = 0.0, x1 = 0.0;
  if (ok) {
    const double2 xv = *reinterpret_cast<const double2*>(a.X + ((size_t)z * n + gr) * NB2 + cq);
    const double2 vv = *reinterpret_cast<const double2*>(a.VP + ((size_t)z * n + gr) * NB2 + cq);
    x0 = xv.x; x1 = xv.y;
    Vs[row][cq] = vv.x; Vs[row][cq + 1] = vv.y;
  } else {
    Vs[row][cq] = 0.0; Vs[row][cq + 1] = 0.0;
  }
  __syncthreads();
  {
    const int p = t / NB2, c = t % NB2;
    double m = 0.0;
    for (int k = 0; k <= p; ++k) m = fma(Ts[k][p], Ss[k][c], m);       // (T^T S)[p][c]
    Ms[p][c] = 0.5 * m;
  }
  __syncthreads();
#pragma unroll 8
  for (int p = 0; p < NB2; ++p) {
    const double v = Vs[row][p];
    x0 = fma(-v, Ms[p][cq], x0);
    x1 = fma(-v, Ms[p][cq + 1], x1);
  }
  if (ok) {
    double* z1 = a.Z1 + ((size_t)z * n + gr) * 2 * NB2;
    double* z2 = a.Z2 + ((size_t)z * n + gr) * 2 * NB2;
    const double2 vv = make_double2(Vs[row][cq], Vs[row][cq + 1]), ww = make_double2(x0, x1);
    *reinterpret_cast<double2*>(z1 + cq) = vv;
    *reinterpret_cast<double2*>(z1 + NB2 + cq) = ww;
    *reinterpret_cast<double2*>(z2 + cq) = ww;
    *reinterpret_cast<double2*>(z2 + NB2 + cq) = vv;
  }
  if (!a.next_cols) return;
  Ws[row][cq] = ok ? x0 : 0.0;
  Ws[row][cq + 1] = ok ? x1 : 0.0;
  {
    const int i = t / NB2, c = t % NB2;          // (S and T are dead: their storage now holds W0 and V0)
    V0s[i][c] = (a.r + i < n) ? a.VP[((size_t)z * n + a.r + i) * NB2 + c] : 0.0;
  }
  __syncthreads();
  {   // W0 = X0 - V0 M  (first NB2 rows of W; same arithmetic as above)
    const int i = t / NB2, c = t % NB2;
    double w = (a.r + i < n) ? a.X[((size_t)z * n + a.r + i) * NB2 + c] : 0.0;
#pragma unroll 8
    for (int p = 0; p < NB2; ++p) w = fma(-V0s[i][p], Ms[p][c], w);
    W0s[i][c] = w;
  }
  __syncthreads();
  if (ok) {
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
    for (int c = 0; c < NB2; ++c) {
      const double v = Vs[row][c], w = Ws[row][c];
      acc0 = fma(v, W0s[cq][c], fma(w, V0s[cq][c], acc0));
      acc1 = fma(v, W0s[cq + 1][c], fma(w, V0s[cq + 1][c], acc1));
    }
    double* cp = a.Cm + (size_t)z * n * a.ldn + (size_t)gr * a.ldn + a.r + cq;
    double2 cv = *reinterpret_cast<double2*>(cp);
    cv.x -= acc0; cv.y -= acc1;
    *reinterpret_cast<double2*>(cp) = cv;
  }
}

// ------------------------------------------------------------------------------------------------
// Band storage for stage 2, row-major: AB[i][e] = C[i][i - e], e = 0 .. 2 NB2 - 1 (row i to the left of its diagonal
// element: the band for e <= NB2, room for the bulge beyond), zero outside the matrix.
__global__ void sb_extract_band_kernel(const double* __restrict__ Cm, double* __restrict__ AB, int n, int ldn) {
  const int z = blockIdx.y;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * LDB) return;
  const int i = (int)(t / LDB), e = (int)(t % LDB);
  double v = 0.0;
  if (e <= NB2 && i - e >= 0) v = Cm[(size_t)z * n * ldn + (size_t)i * ldn + (i - e)];
  AB[(size_t)z * n * LDB + t] = v;
}

__global__ void sb_extract_tridiag_kernel(const double* __restrict__ AB, double* __restrict__ dd, double* __restrict__ ee,
                                          int n) {
  const int z = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  dd[(size_t)z * n + j] = AB[((size_t)z * n + j) * LDB];
  ee[(size_t)z * n + j] = (j + 1 < n) ? AB[((size_t)z * n + j + 1) * LDB + 1] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Stage 2: bulge chasing.  Sweep s annihilates column s of the band below the sub-diagonal with a reflector on
// rows [s+1, s+1+NB2) and chases the bulge down the band in steps of NB2 rows; step k >= 1 of sweep s works on the
// rows [q0, q0+NB2), q0 = s+1+k NB2: the off-diagonal block B = A[q0:, r0:q0] (r0 = q0 - NB2) and the diagonal
// block D = A[q0:, q0:]:
//     B <- B H_prev;  new reflector H' from B[:, 0];  B[:, 1:] <- H' B[:, 1:];  D <- H' D H'.
// Step (s+1, k) may run once (s, k+1) is done, so consecutive sweeps follow each other two steps apart and the
// critical path is 2 n dependent steps: the step latency is everything.  A CTA therefore owns a GROUP of CGW
// consecutive sweeps, one warp per sweep, running as a wavefront (warp w is at step tau - 2 w in time step tau) on
// a sliding window of band ROWS kept in shared memory: inside a group a hand-over costs one __syncthreads instead of
// a round trip through L2.  In one time step the warps touch disjoint rows.  A fifth warp streams the window: it
// loads the 32 rows the leading sweep needs next (after the previous group has published them), writes back the
// rows the trailing sweep has finished and publishes the group's row bound (release / acquire counters in global
// memory); both overlap the compute warps' step.  One warp per step: lane = row of the blocks (32 + 32 doubles per
// lane in registers), vectors are broadcast through a per-warp shared line, column sums use a butterfly
// transpose-reduction.
constexpr int CGW = 3;                        // sweeps per group
constexpr int CG_NSLOT = 256;                 // window rows (power of two)
constexpr int CG_PITCH = LDB + 2;             // 66: 16-byte aligned rows, lane stride 67 -> conflict-free block access
constexpr int CG_THREADS = (2 * CGW + 2) * 32;   // two warps per sweep + storer + fetcher
static_assert(63 * CGW - 31 + NB2 <= CG_NSLOT, "window too small");   // live rows of a time step + the chunk in flight

__device__ __forceinline__ void bcast_store(double* line, double v, int lane) {
  __syncwarp();
  line[lane] = v;
  __syncwarp();
}

// lane c receives sum over lanes of vals[c] (vals is destroyed)
__device__ __forceinline__ double colsum32(double (&vals)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double send = up ? vals[i] : vals[i + o];
      const double keep = up ? vals[i + o] : vals[i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return vals[0];
}

// Householder generator over the lanes (LAPACK dlarfg): x_lane -> v_lane (v_0 = 1), tau, beta.
__device__ __forceinline__ void warp_house(double x, int lane, double& v, double& tau, double& beta) {
  const double alpha = __shfl_sync(0xffffffffu, x, 0);
  const double ss = warp_sum(lane > 0 ? x * x : 0.0);
  if (ss == 0.0) {
    beta = alpha; tau = 0.0; v = (lane == 0) ? 1.0 : 0.0;
  } else {
    // one reciprocal square root and one reciprocal instead of a square root and two divisions (this chain sits on
    // the critical path of every chase step):  |beta| = nrm2 r,  tau = (beta - alpha) / beta = 1 + |alpha| r
    const double nrm2 = fma(alpha, alpha, ss), r = rsqrt(nrm2);
    beta = -copysign(nrm2 * r, alpha);
    tau = fma(fabs(alpha), r, 1.0);
    v = (lane == 0) ? 1.0 : x * (1.0 / (alpha - beta));
  }
}

// D <- H D H for the symmetric block whose row `lane` is in D[] (full rows); v, tau = reflector; line = 2 x 32 doubles.
__device__ __forceinline__ void two_sided32(double (&D)[32], double v, double tau, int lane, double* line) {
  bcast_store(line, v, lane);
  double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
  for (int c = 0; c < 32; c += 4) {
    const double2 va = *reinterpret_cast<const double2*>(line + c);
    const double2 vb = *reinterpret_cast<const double2*>(line + c + 2);
    p0 = fma(D[c], va.x, p0);
    p1 = fma(D[c + 1], va.y, p1);
    p2 = fma(D[c + 2], vb.x, p2);
    p3 = fma(D[c + 3], vb.y, p3);
  }
  const double p = tau * ((p0 + p1) + (p2 + p3));
  const double gamma = 0.5 * tau * warp_sum(p * v);
  const double w = p - gamma * v;
  line[32 + lane] = w;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double2 vv = *reinterpret_cast<const double2*>(line + c);
    const double2 ww = *reinterpret_cast<const double2*>(line + 32 + c);
    D[c] -= v * ww.x + w * vv.x;
    D[c + 1] -= v * ww.y + w * vv.y;
  }
}

struct SbChase {
  double* AB; double* V2; int* gprog;
  int n, ldn, nz;
  long long* dbg;      // APV_TS_DEBUG: clock64 totals of CTA 0 (compute step / barrier wait; loader store / wait / load)
};
#define CH_TICK(k) do { if (dbgp) { const long long _t = clock64(); if (lane == 0) dbgp[k] += _t - tk; tk = _t; } } while (0)
// phases inside a step (compile with -DAPV_CHASE_FINE: the checks cost 1 ms of the kernel's 19.5 even when switched off)
#ifdef APV_CHASE_FINE
#define CH_FINE(k) do { if (fine) { const long long _t = clock64(); if (lane == 0) fine[k] += _t - tf; tf = _t; } } while (0)
#define CH_FINE_START() do { if (fine) tf = clock64(); } while (0)
#else
#define CH_FINE(k)
#define CH_FINE_START()
#endif

__device__ __forceinline__ int sweep_steps(int n, int s) { return (s <= n - 3) ? 1 + (n - s - 2) / NB2 : 0; }

__device__ __forceinline__ double* win_row(double* win, int i) { return win + (size_t)(i & (CG_NSLOT - 1)) * CG_PITCH; }

// Row `lane` of the diagonal block at q0 from the window: own row to the left of the diagonal, the rest by symmetry.
__device__ __forceinline__ void win_load_diag(double* win, int q0, int lane, double (&D)[32]) {
  const double* pd = win_row(win, q0 + lane) + lane;
#pragma unroll
  for (int c = 0; c < 32; ++c) D[c] = (c <= lane) ? pd[-c] : win_row(win, q0 + c)[c - lane];
}

__device__ __forceinline__ void win_store_diag(double* win, int q0, int lane, const double (&D)[32]) {
  double* pd = win_row(win, q0 + lane) + lane;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane) pd[-c] = D[c];
}

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(CG_THREADS, 1) sb2st_chase_kernel(SbChase a) {
  extern __shared__ __align__(16) double win[];                  // [CG_NSLOT][CG_PITCH] rows of the band
  __shared__ __align__(16) double lines[2 * CGW][64];            // per-warp broadcast lines
  __shared__ __align__(16) double vline[CGW][34];                // new reflector of the step: v[32], tau
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // roles: warps 2w / 2w+1 = off-diagonal-block / diagonal-block half of sweep w's step, then storer and fetcher
  const int role = wib < 2 * CGW ? (wib & 1) : 2 + (wib - 2 * CGW);
  const int wsw = wib >> 1;
  const int nz = a.nz, z = blockIdx.x % nz, n = a.n;
  const int cta = blockIdx.x / nz, G = gridDim.x / nz;
  double* AB = a.AB + (size_t)z * n * LDB;
  double* V2 = a.V2 + (size_t)z * n * a.ldn;
  const int ngroups = (n - 2 + CGW - 1) / CGW;
  int* gprog = a.gprog + (size_t)z * ngroups;
  long long* dbgp = nullptr;
#ifdef APV_CHASE_FINE
  long long* fine = nullptr;              // phases inside a step: B warp of sweep 0 -> dbg[20..27], its D warp -> dbg[28..31]
  long long tf = clock64();
#endif
  long long tk = clock64();

  // fetcher: rows [i0, i0 + 32) of the band -> window (zero rows beyond the matrix) once the previous group is past them
  auto load_chunk = [&](int g, int i0) {
    if (g > 0 && i0 < n) {
      const int need = min(n, i0 + NB2);
      if (lane == 0) {
        // relaxed polling with back-off, one acquire fence when the bound is reached
        int v;
        for (;;) {
          asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(gprog + (g - 1)) : "memory");
          if (v >= need) break;
          __nanosleep(32);
        }
        __threadfence();
      }
      __syncwarp();
    }
    CH_TICK(1);
    double2 buf[NB2];                   // all 32 rows in flight at once: one L2 round trip per chunk
#pragma unroll
    for (int u = 0; u < NB2; ++u) {
      const int i = i0 + u;
      buf[u] = (i < n) ? __ldcg(reinterpret_cast<const double2*>(AB + (size_t)i * LDB) + lane) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < NB2; ++u) reinterpret_cast<double2*>(win_row(win, i0 + u))[lane] = buf[u];
    CH_TICK(2);
  };
  // storer: window rows [y0, y1) -> global, then publish the bound
  auto store_rows = [&](int g, int y0, int y1, int publish) {
    for (int i = y0; i < y1; i += 8) {
      double2 buf[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + u < y1) buf[u] = reinterpret_cast<const double2*>(win_row(win, i + u))[lane];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + u < y1) reinterpret_cast<double2*>(AB + (size_t)(i + u) * LDB)[lane] = buf[u];
    }
    // the warp barrier orders the lanes' stores before lane 0's release (release is cumulative)
    __syncwarp();
    if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gprog + g), "r"(publish) : "memory");
    CH_TICK(0);
  };

  for (int g = cta; g < ngroups; g += G) {
    // debug clocks: the first group only (it never waits for a predecessor: the unthrottled pace of one time step)
    dbgp = (a.dbg && blockIdx.x == 0 && g == 0 && (wib == 0 || role >= 2)) ? a.dbg + 8 + (role >= 2 ? 4 * (role - 1) : 0) : nullptr;
#ifdef APV_CHASE_FINE
    fine = (a.dbg && blockIdx.x == 0 && g == 0 && wib < 2) ? a.dbg + (wib == 0 ? 20 : 28) : nullptr;
#endif
    tk = clock64();
    const int s0 = g * CGW;
    const int nst0 = sweep_steps(n, s0);
    int T = 0;
#pragma unroll
    for (int w = 0; w < CGW; ++w) {
      const int ns = sweep_steps(n, s0 + w);
      if (ns > 0) T = max(T, ns + 2 * w);
    }
    if (role == 3) load_chunk(g, s0 + 1);
    __syncthreads();
    const int s = s0 + wsw;
    const int nst = (role < 2) ? sweep_steps(n, s) : 0;
    double v = 0.0, tau = 0.0, beta = 0.0;      // role 0: reflector carried from step to step
    int ystored = s0 + 1;                       // storer: rows below are back in global memory
    for (int t = 0; t < T; ++t) {
      if (role < 2) {
        const int k = t - 2 * wsw;
        if (k >= 0 && k < nst) {
          const int q0 = s + 1 + k * NB2;       // rows of the step (k = 0: the first diagonal block)
          double* line = lines[wib];
          CH_FINE_START();
          if (role == 0) {
            double* myrow = win_row(win, q0 + lane);
            if (k == 0) {
              // reflector of column s
              warp_house(myrow[1 + lane], lane, v, tau, beta);          // A[s+1+lane][s]
              vline[wsw][lane] = v;
              if (lane == 0) { vline[wsw][32] = tau; myrow[1] = beta; }
              pair_sync(1 + wsw);
            } else {
              double B[32];
              double* pb = myrow + NB2 + lane;                        // B(lane, c) = pb[-c]
#pragma unroll
              for (int c = 0; c < 32; ++c) B[c] = pb[-c];
              CH_FINE(0);
              // B <- B H_prev
              bcast_store(line, v, lane);
              double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                const double2 va = *reinterpret_cast<const double2*>(line + c);
                const double2 vb = *reinterpret_cast<const double2*>(line + c + 2);
                y0 = fma(B[c], va.x, y0);
                y1 = fma(B[c + 1], va.y, y1);
                y2 = fma(B[c + 2], vb.x, y2);
                y3 = fma(B[c + 3], vb.y, y3);
              }
              const double y = tau * ((y0 + y1) + (y2 + y3));
              // the first column decides the new reflector: finish it first and hand it to the diagonal-block warp
              B[0] = fma(-y, line[0], B[0]);
              CH_FINE(1);
              double vn, taun;
              warp_house(B[0], lane, vn, taun, beta);
              CH_FINE(2);
              vline[wsw][lane] = vn;
              if (lane == 0) vline[wsw][32] = taun;
              pair_sync(1 + wsw);
              CH_FINE(3);
#pragma unroll
              for (int c = 2; c < 32; c += 2) {
                const double2 vv = *reinterpret_cast<const double2*>(line + c);
                B[c] = fma(-y, vv.x, B[c]);
                B[c + 1] = fma(-y, vv.y, B[c + 1]);
              }
              B[1] = fma(-y, line[1], B[1]);
              v = vn; tau = taun;
              B[0] = (lane == 0) ? beta : 0.0;
              CH_FINE(4);
              // B[:, 1:] <- H' B[:, 1:]
              double vals[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) vals[c] = v * B[c];
              const double u = tau * colsum32(vals, lane);       // lane c: tau v^T B[:, c]
              CH_FINE(5);
              bcast_store(line, u, lane);
#pragma unroll
              for (int c = 2; c < 32; c += 2) {
                const double2 uu = *reinterpret_cast<const double2*>(line + c);
                B[c] = fma(-v, uu.x, B[c]);
                B[c + 1] = fma(-v, uu.y, B[c + 1]);
              }
              B[1] = fma(-v, line[1], B[1]);
              CH_FINE(6);
#pragma unroll
              for (int c = 0; c < 32; ++c) pb[-c] = B[c];
              CH_FINE(7);
            }
            if (q0 + lane < n) V2[(size_t)s * a.ldn + q0 + lane] = (lane == 0) ? tau : v;
          } else {
            // diagonal block: D <- H' D H'
            double D[32];
            win_load_diag(win, q0, lane, D);
            CH_FINE(0);
            pair_sync(1 + wsw);
            CH_FINE(1);
            two_sided32(D, vline[wsw][lane], vline[wsw][32], lane, line);
            CH_FINE(2);
            win_store_diag(win, q0, lane, D);
            CH_FINE(3);
          }
        }
        CH_TICK(0);
      } else if (role == 2) {
        // ---- storer: write back what time step t - 1 finished and publish the group's row bound
        if (t > 0) {
          int ynew = n, remaining = 0;
#pragma unroll
          for (int w = 0; w < CGW; ++w) {
            const int kdone = t - 1 - 2 * w, ns = sweep_steps(n, s0 + w);
            if (ns > 0 && kdone + 1 < ns) {
              remaining = 1;
              ynew = min(ynew, s0 + w + 1 + NB2 * max(kdone + 1, 0));
            }
          }
          ynew = min(ynew, n);
          if (ynew > ystored) {
            store_rows(g, ystored, ynew, remaining ? ynew : PROG_DONE);
            ystored = ynew;
          }
        }
      } else {
        // ---- fetcher: the rows the leading sweep needs in time step t + 1
        if (t + 1 <= nst0) load_chunk(g, s0 + 1 + (t + 1) * NB2);
      }
      __syncthreads();
      CH_TICK(3);
    }
    if (role == 2) store_rows(g, ystored, min(n, s0 + 1 + (nst0 + 1) * NB2), PROG_DONE);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// z <- Q2 z for the V eigenvectors of T: sweeps in descending order (the reflectors of one sweep act on disjoint
// rows), one CTA per vector, warp w applies the reflectors k = w, w + 32, .. of the sweep.  grid (V, nz), 1024 thr.
constexpr int Q2T = 1024;
constexpr int Q2R = 8;            // reflectors per warp and sweep: n <= 32 * 32 * Q2R = 8192

__global__ void __launch_bounds__(Q2T) sb2st_apply_q2_kernel(double* __restrict__ iv, const double* __restrict__ V2,
                                                             int n, int ldn, int Vp) {
  extern __shared__ double xs[];
  const int v = blockIdx.x, z = blockIdx.y;
  double* X = iv + (size_t)z * 6 * n * Vp + 4 * (size_t)n * Vp + v;
  for (int i = threadIdx.x; i < n; i += Q2T) xs[i] = X[(size_t)i * Vp];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* v2 = V2 + (size_t)z * n * ldn;
  double cur[Q2R], nxt[Q2R];
  auto fetch = [&](int s, double (&buf)[Q2R]) {
#pragma unroll
    for (int u = 0; u < Q2R; ++u) {
      const int row = s + 1 + NB2 * (warp + 32 * u) + lane;
      buf[u] = (s >= 0 && row < n) ? __ldg(v2 + (size_t)s * ldn + row) : 0.0;
    }
  };
  fetch(n - 3, cur);
  for (int s = n - 3; s >= 0; --s) {
    fetch(s - 1, nxt);
#pragma unroll
    for (int u = 0; u < Q2R; ++u) {
      const int row = s + 1 + NB2 * (warp + 32 * u) + lane;
      if (s + 1 + NB2 * (warp + 32 * u) < n) {          // warp-uniform
        const double tau = __shfl_sync(0xffffffffu, cur[u], 0);
        const double vi = (lane == 0) ? 1.0 : cur[u];
        const double x = (row < n) ? xs[row] : 0.0;
        const double dot = warp_sum(vi * x);
        if (row < n) xs[row] = fma(-tau * dot, vi, x);
      }
    }
#pragma unroll
    for (int u = 0; u < Q2R; ++u) cur[u] = nxt[u];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += Q2T) X[(size_t)i * Vp] = xs[i];
}


// ------------------------------------------------------------------------------------------------
// Z <- Q2 Z, wavefront form: lane = eigenvector (groups of 32 vectors are independent problems), so a reflector is
// applied with thread-local dot products -- no shuffles, no CTA barriers.  A warp owns a CHAIN: the 32 consecutive
// sweeps [S, S+32) for 32 vectors, and walks down the band in blocks k = 0, 1, ..: block k applies the reflectors
// (s, k), s = S+31 .. S, which act on the rows [S+1+32k, S+64+32k) -- a window of 63 rows per lane in registers
// that slides by 32 rows per block.  Chain S may run block k once chain S+32 has finished block k (the reflectors of
// the higher sweeps come first); finished rows travel between chains through global memory with release /
// acquire block counters, like the rows of the bulge chasing.  The reflectors of a block (32 x 32 doubles) are
// staged in shared memory and read as broadcasts.  Cooperative launch, 128 threads, one chain per warp.
constexpr int Q2W_T = 256;        // (8 chains per CTA: the same chains in flight on half as many SMs as with 4)
constexpr int Q2W_MAXV = 8192;

struct SbQ2 {
  double* iv; const double* V2; int* prog;
  int n, ldn, Vp, nz, ngrp, nhalf;
};

__global__ void __launch_bounds__(Q2W_T, 1) sb2st_apply_q2_wave_kernel(SbQ2 a) {
  extern __shared__ __align__(16) double q2w_sm[];                 // [Q2W_T / 32][32][34]
  double (*vs_all)[32][34] = reinterpret_cast<double (*)[32][34]>(q2w_sm);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = a.n, ldn = a.ldn, Vp = a.Vp;
  const int nprob = a.nz * a.nhalf;
  const int gw = blockIdx.x * (Q2W_T / 32) + wib, GW = gridDim.x * (Q2W_T / 32);
  const int prob = gw % nprob, wp = gw / nprob, Gp = GW / nprob;      // warps are dealt round-robin to the problems
  if (wp >= Gp) return;
  const int z = prob / a.nhalf, h = prob % a.nhalf;
  double* X = a.iv + (size_t)z * 6 * n * Vp + 4 * (size_t)n * Vp + 32 * h + lane;       // X[i * Vp]
  const double* v2 = a.V2 + (size_t)z * n * ldn;
  int* prog = a.prog + (size_t)prob * a.ngrp;
  double (*vs)[34] = vs_all[wib];
  for (int c = wp; c < a.ngrp; c += Gp) {          // chains in descending sweep order
    const int gi = a.ngrp - 1 - c, S = gi * NB2;
    const int K = 1 + (n - S - 2) / NB2;            // blocks of this chain (= steps of its first sweep)
    double w[63];
    for (int k = 0; k < K; ++k) {
      const int base = S + 1 + k * NB2;             // first row of the window
      if (gi + 1 < a.ngrp) {                        // the chain above must have finished block k
        if (lane == 0) {
          int v;
          for (;;) {
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog + gi + 1) : "memory");
            if (v >= k + 1) break;
            __nanosleep(32);
          }
          __threadfence();
        }
        __syncwarp();
      }
      // stage the block's reflectors: vs[o][i] = V2[S + o][base + o + i]  (entry 0 holds tau); zero beyond the matrix
#pragma unroll 8
      for (int o = 0; o < 32; ++o) {
        const int s = S + o, row = base + o + lane;
        vs[o][lane] = (s <= n - 3 && row < n) ? __ldcg(v2 + (size_t)s * ldn + row) : 0.0;
      }
      // rows entering the window
      if (k == 0) {
#pragma unroll
        for (int j = 0; j < 63; ++j) w[j] = (base + j < n) ? __ldcg(X + (size_t)(base + j) * Vp) : 0.0;
      } else {
#pragma unroll
        for (int j = 31; j < 63; ++j) w[j] = (base + j < n) ? __ldcg(X + (size_t)(base + j) * Vp) : 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int o = 31; o >= 0; --o) {
        if (S + o <= n - 3 && base + o < n) {        // warp-uniform: the reflector exists
          const double tau = vs[o][0];
          double d0 = w[o], d1 = 0.0;                // v_0 = 1
#pragma unroll
          for (int i = 2; i < 32; i += 2) {
            const double2 vv = *reinterpret_cast<const double2*>(&vs[o][i]);
            d0 = fma(vv.x, w[o + i], d0);
            d1 = fma(vv.y, w[o + i + 1], d1);
          }
          d1 = fma(vs[o][1], w[o + 1], d1);
          const double td = tau * (d0 + d1);
          w[o] -= td;
          w[o + 1] = fma(-td, vs[o][1], w[o + 1]);
#pragma unroll
          for (int i = 2; i < 32; i += 2) {
            const double2 vv = *reinterpret_cast<const double2*>(&vs[o][i]);
            w[o + i] = fma(-td, vv.x, w[o + i]);
            w[o + i + 1] = fma(-td, vv.y, w[o + i + 1]);
          }
        }
      }
      // rows leaving the window are final for this chain (all of it after the last block)
      const int nout = (k + 1 == K) ? 63 : 32;
#pragma unroll
      for (int j = 0; j < 63; ++j)
        if (j < nout && base + j < n) X[(size_t)(base + j) * Vp] = w[j];
#pragma unroll
      for (int j = 0; j < 31; ++j) w[j] = w[j + 32];
      __syncwarp();
      if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(prog + gi), "r"(k + 1 == K ? PROG_DONE : k + 1) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Z <- Q1 Z with the stage-1 block reflectors  Q1 = P_0 P_1 ..,  P_p = I - V_p T_p V_p^T  (the compact-WY factors of
// the panel QR), panels in descending order, for a chunk of Q1VC eigenvectors at a time.  The reflector panels are
// streamed ONCE: a CTA owns a slab of Q1RS rows of all the vectors of the chunk (kept in shared memory for the whole
// kernel) and per panel forms its part of S = V^T Z; one grid barrier per zone; every CTA sums the partial S, applies
// T and updates its rows.  (One CTA per vector re-read every panel from L2: 17 GB of L2 traffic at n = 4096.)
// Cooperative launch, grid = (n / Q1RS) CTAs per zone, 512 threads.
constexpr int Q1T = 512;
constexpr int Q1RS = 128;        // rows per CTA (n = 8192, two zones: 128 CTAs, all co-resident)
constexpr int Q1VC = 64;         // vectors per chunk
constexpr int Q1VP = NB2 + 2;    // pitch of the staged reflector rows (even: 16-byte loads, conflict-free)

struct SbQ1 {
  const double* iv; const double* VH; const double* Tall; double* Zt; double* Sp; unsigned long long* bar;
  int n, ldn, V, Vp, nz, npanels, v0;
};

__device__ __forceinline__ void q1_zone_sync(unsigned long long* bar, int G) {
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

__global__ void __launch_bounds__(Q1T, 1) sb_apply_q1_kernel(SbQ1 a) {
  extern __shared__ __align__(16) double q1sm[];
  double* Zs = q1sm;                               // [Q1RS][Q1VC]   the slab of the vectors
  double* Vs = Zs + Q1RS * Q1VC;                   // [Q1RS][Q1VP]   the slab of the current reflector panel
  double* Us = Vs + Q1RS * Q1VP;                   // [NB2][Q1VC]    T V^T Z
  const int nz = a.nz, z = blockIdx.x % nz, g = blockIdx.x / nz, G = gridDim.x / nz;
  const int n = a.n, ldn = a.ldn, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R0 = g * Q1RS, nr = max(0, min(Q1RS, n - R0));
  const int nv = min(Q1VC, a.V - a.v0);
  const double* X = a.iv + (size_t)z * 6 * n * a.Vp + 4 * (size_t)n * a.Vp + a.v0;
  for (int i = tid; i < Q1RS * Q1VC; i += Q1T) {
    const int rr = i / Q1VC, v = i % Q1VC;
    Zs[i] = (rr < nr && v < nv) ? X[(size_t)(R0 + rr) * a.Vp + v] : 0.0;
  }
  const double* vh = a.VH + (size_t)z * n * ldn;
  double* Sp = a.Sp + (size_t)z * 2 * G * NB2 * Q1VC;
  for (int p = a.npanels - 1; p >= 0; --p) {
    const int j0 = p * NB2, r = j0 + NB2;
    const bool active = R0 + nr > r;               // (the rows above r see zeros of V)
    double* spg = Sp + ((size_t)(p & 1) * G + g) * NB2 * Q1VC;
    __syncthreads();
    if (active) {
      // stage the panel rows of the slab: Vs[row][c] = v_{j0+c}[R0 + row]
      for (int i = tid; i < NB2 * Q1RS; i += Q1T) {
        const int c = i / Q1RS, rr = i % Q1RS;
        Vs[rr * Q1VP + c] = (rr < nr) ? __ldg(vh + (size_t)(j0 + c) * ldn + R0 + rr) : 0.0;
      }
      __syncthreads();
      // partial S[c][4 vq .. 4 vq + 3], c = lane, vq = warp
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const int rlo = max(0, r - R0);
#pragma unroll 8
      for (int rr = rlo; rr < nr; ++rr) {
        const double vv = Vs[rr * Q1VP + lane];
        const double2 za = *reinterpret_cast<const double2*>(Zs + rr * Q1VC + 4 * warp);
        const double2 zb = *reinterpret_cast<const double2*>(Zs + rr * Q1VC + 4 * warp + 2);
        s0 = fma(vv, za.x, s0); s1 = fma(vv, za.y, s1); s2 = fma(vv, zb.x, s2); s3 = fma(vv, zb.y, s3);
      }
      double* o = spg + lane * Q1VC + 4 * warp;
      o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3;
    } else {
      for (int i = tid; i < NB2 * Q1VC; i += Q1T) spg[i] = 0.0;
    }
    q1_zone_sync(a.bar + z, G);
    if (!active) continue;
    // S = sum of the partials, U = T S  (thread: row c = tid / 16 of U, vectors 4 (tid % 16) ..)
    {
      const int c = tid >> 4, v4 = (tid & 15) * 4;
      // the summed S goes through shared memory (Us), then U = T S overwrites it
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
      const double* sp = Sp + (size_t)(p & 1) * G * NB2 * Q1VC + c * Q1VC + v4;
#pragma unroll 8
      for (int q = 0; q < G; ++q) {
        const double2 x0 = __ldcg(reinterpret_cast<const double2*>(sp + (size_t)q * NB2 * Q1VC));
        const double2 x1 = __ldcg(reinterpret_cast<const double2*>(sp + (size_t)q * NB2 * Q1VC + 2));
        t0 += x0.x; t1 += x0.y; t2 += x1.x; t3 += x1.y;
      }
      Us[c * Q1VC + v4] = t0; Us[c * Q1VC + v4 + 1] = t1; Us[c * Q1VC + v4 + 2] = t2; Us[c * Q1VC + v4 + 3] = t3;
      __syncthreads();
      const double* T = a.Tall + ((size_t)p * nz + z) * NB2 * NB2 + (size_t)c * NB2;       // row c of T (upper)
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll 8
      for (int d = c; d < NB2; ++d) {
        const double td = __ldg(T + d);
        u0 = fma(td, Us[d * Q1VC + v4], u0); u1 = fma(td, Us[d * Q1VC + v4 + 1], u1);
        u2 = fma(td, Us[d * Q1VC + v4 + 2], u2); u3 = fma(td, Us[d * Q1VC + v4 + 3], u3);
      }
      __syncthreads();
      Us[c * Q1VC + v4] = u0; Us[c * Q1VC + v4 + 1] = u1; Us[c * Q1VC + v4 + 2] = u2; Us[c * Q1VC + v4 + 3] = u3;
      __syncthreads();
    }
    // Z[row][v] -= sum_c V[row][c] U[c][v]:  thread: vector v = tid % 64 (U column in registers), rows tid / 64 + 8 k
    {
      const int v = tid & 63, rg = tid >> 6;
      double u[NB2];
#pragma unroll
      for (int c = 0; c < NB2; ++c) u[c] = Us[c * Q1VC + v];
      const int rlo = max(0, r - R0);
#pragma unroll 2
      for (int rr = rlo + rg; rr < nr; rr += Q1T / 64) {
        const double2* vr = reinterpret_cast<const double2*>(Vs + rr * Q1VP);
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int c = 0; c < NB2; c += 2) {
          const double2 vv = vr[c >> 1];
          acc0 = fma(vv.x, u[c], acc0);
          acc1 = fma(vv.y, u[c + 1], acc1);
        }
        Zs[rr * Q1VC + v] -= acc0 + acc1;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nv * Q1RS; i += Q1T) {
    const int v = i / Q1RS, rr = i % Q1RS;
    if (rr < nr) a.Zt[((size_t)z * a.V + a.v0 + v) * n + R0 + rr] = Zs[rr * Q1VC + v];
  }
}

}  // namespace

size_t twostage_scratch_bytes(int n, int nz, int nsplit_max) {
  size_t d = 0;
  d += (size_t)nz * n * NB2;                          // VP
  d += (size_t)nz * nsplit_max * n * NB2;             // Ypart
  d += (size_t)nz * n * NB2;                          // X
  d += (size_t)nz * ceil_div(n, 64) * NB2 * NB2;      // Spart
  d += (size_t)nz * ceil_div(n, NB2) * NB2 * NB2;     // T factors of all panels [panel][zone][NB2][NB2]
  d += (size_t)nz * n * LDB;                          // AB
  return d * sizeof(double) + (size_t)nz * n * sizeof(int);
}

int twostage_nsplit_max() { return 8; }

// APV_TS_DEBUG=1: synchronous per-kernel-class timing of the two-stage reduction, printed to stderr.
struct TsTrace {        // APV_TS_DEBUG=2: events on the main stream around every kernel class, read after the loop
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cls;
  size_t used = 0;
  void mark(cudaStream_t st, int c) {
    if (!on) return;
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); cls.push_back(0); }
    cls[used] = c;
    cudaEventRecord(ev[used++], st);
  }
};

struct TsDebug {
  bool on = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void begin(cudaStream_t st) { if (on) cudaEventRecord(e0, st); }
  void end(cudaStream_t st, int k) {
    if (!on) return;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    acc[k] += ms;
  }
};

int twostage_run(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, nz = ws.nz;
  TsDebug dbg;
  static long long* dclk = nullptr;
  if (getenv("APV_TS_DEBUG")) {
    dbg.on = true;
    cudaEventCreate(&dbg.e0);
    cudaEventCreate(&dbg.e1);
    if (!dclk) cudaMalloc((void**)&dclk, 32 * sizeof(long long));
    cudaMemsetAsync(dclk, 0, 32 * sizeof(long long), st);
  }
  const long long mstride = (long long)n * ldn;
  const int nsm = twostage_nsplit_max();
  double* VP = ws.ts2;
  double* Ypart = VP + (size_t)nz * n * NB2;
  double* X = Ypart + (size_t)nz * nsm * n * NB2;
  double* Spart = X + (size_t)nz * n * NB2;
  double* Tp = Spart + (size_t)nz * ceil_div(n, 64) * NB2 * NB2;
  double* AB = Tp + (size_t)nz * ceil_div(n, NB2) * NB2 * NB2;
  int* prog = reinterpret_cast<int*>(AB + (size_t)nz * n * LDB);
  int dev = 0, sms = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  APV_CUDA_TRY(cudaMemsetAsync(ws.tau, 0, (size_t)nz * n * sizeof(double), st));
  APV_CUDA_TRY(cudaMemsetAsync(prog, 0, (size_t)nz * n * sizeof(int), st));

  // ---- stage 1
  // Look-ahead: as soon as panel k's W is known, the NB2 columns of the next panel are updated on their own (inside
  // sb_w2_kernel), the next panel's QR (latency-bound, one cluster per zone) starts, and the bulk of the rank-2b
  // update (all the other SMs) runs beside it on a side stream.
  static TsTrace tr;
  tr.on = getenv("APV_TS_DEBUG") && atoi(getenv("APV_TS_DEBUG")) == 2;
  tr.used = 0;
  if (tr.on) dbg.on = false;
  const bool lookahead = !dbg.on && ws.st2 != nullptr && !getenv("APV_TS_NO_LOOKAHEAD");
  auto launch_qr = [&](int j0, cudaStream_t qs) -> int {
    const int r = j0 + NB2, npn = n - r;
    // cluster size: smallest of 1, 2, 4, 8, 16 whose slab fits in the registers of a CTA (16 rows per warp) and in its
    // shared memory (output staging)
    const int max_rows = std::min(16 * QRW, (184 * 1024) / (PP * (int)sizeof(double)));
    int CS = 1;
    while (CS < QR_MAXCS && ceil_div(npn, CS) > max_rows) CS *= 2;
    if (npn > 64) CS = std::max(CS, 8);                       // spread the rows anyway: the column loop is latency-bound
    static const int cs16_from = getenv("APV_QR_CS16_FROM") ? atoi(getenv("APV_QR_CS16_FROM")) : 0;   // (experiment)
    if (cs16_from > 0 && npn >= cs16_from) CS = std::max(CS, 16);
    if (ceil_div(npn, CS) > max_rows) {
      snprintf(g_err, sizeof(g_err), "two-stage tridiagonalisation: n = %d exceeds the panel capacity", n);
      return EINVAL_;
    }
    SbPanel p;
    p.Cm = ws.Cm; p.VH = ws.VH; p.VP = VP; p.tau = ws.tau; p.Tp = Tp + (size_t)(j0 / NB2) * nz * NB2 * NB2;
    p.n = n; p.ldn = ldn; p.j0 = j0; p.rows_per = ceil_div(npn, CS); p.dbg = dbg.on ? dclk : nullptr;
    const size_t smem = (size_t)p.rows_per * PP * sizeof(double);
    static PerDevice pd_configured; size_t& configured = pd_configured.cur();
    if (smem > configured) {
      const int want = (int)std::max(smem, (size_t)(184 * 1024));
      for (const void* f : {(const void*)sb_panel_qr_kernel<2>, (const void*)sb_panel_qr_kernel<4>,
                            (const void*)sb_panel_qr_kernel<8>, (const void*)sb_panel_qr_kernel<16>}) {
        APV_CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
        APV_CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      }
      configured = (size_t)want;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, nz); cfg.blockDim = dim3(QRT); cfg.dynamicSmemBytes = smem; cfg.stream = qs;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    dbg.begin(qs);
    // rows per warp held in registers: the smallest instantiation that covers the slab
    const int rpw = ceil_div(p.rows_per, QRW);
    if (rpw > 16) {
      snprintf(g_err, sizeof(g_err), "two-stage tridiagonalisation: n = %d exceeds the panel capacity", n);
      return EINVAL_;
    }
    if (rpw <= 2) APV_CUDA_TRY(cudaLaunchKernelEx(&cfg, sb_panel_qr_kernel<2>, p));
    else if (rpw <= 4) APV_CUDA_TRY(cudaLaunchKernelEx(&cfg, sb_panel_qr_kernel<4>, p));
    else if (rpw <= 8) APV_CUDA_TRY(cudaLaunchKernelEx(&cfg, sb_panel_qr_kernel<8>, p));
    else APV_CUDA_TRY(cudaLaunchKernelEx(&cfg, sb_panel_qr_kernel<16>, p));
    dbg.end(qs, 0);
    ++*launches;
    return OK;
  };
  bool qr_done = false;          // the QR of the current panel was issued by the previous iteration (look-ahead)
  for (int j0 = 0; n - j0 - NB2 >= 2; j0 += NB2) {
    const int r = j0 + NB2, npn = n - r;
    tr.mark(st, 0);            // [0] join with the look-ahead QR / own QR
    if (!qr_done) APV_TRY(launch_qr(j0, st));
    tr.mark(st, 1);            // [1] Y
    // Y = C22 V in K slices
    int nsplit = std::max(1, std::min(nsm, (2 * sms) / std::max(1, nz * ceil_div(npn, 128))));
    const int kslice = round_up(ceil_div(npn, nsplit), 16);
    nsplit = ceil_div(npn, kslice);
    GemmArgs y{};
    y.batch = nz; y.split = nsplit; y.split_ktot = npn;
    y.A = ws.Cm + (size_t)r * ldn + r; y.lda = ldn; y.strideA = mstride; y.splitA = kslice;
    y.B = VP + (size_t)r * NB2; y.ldb = NB2; y.strideB = (long long)n * NB2; y.splitB = (long long)kslice * NB2;
    y.C = Ypart + (size_t)r * NB2; y.ldc = NB2; y.strideC = (long long)nsplit * n * NB2; y.splitC = (long long)n * NB2;
    y.M = npn; y.N = NB2; y.K = kslice; y.alpha = 1.0; y.beta = 0.0;
    dbg.begin(st);
    APV_TRY(gemm_f64(y, st));
    dbg.end(st, 1);
    ++*launches;
    tr.mark(st, 2);            // [2] W (+ next panel's columns)
    SbW w;
    w.Ypart = Ypart; w.VP = VP; w.Tp = Tp + (size_t)(j0 / NB2) * nz * NB2 * NB2; w.X = X; w.Spart = Spart; w.Z1 = ws.Z1; w.Z2 = ws.Z2;
    w.n = n; w.r = r; w.npn = npn; w.nsplit = nsplit; w.nblk = ceil_div(npn, 64);
    const bool next_panel = n - r - NB2 >= 2;
    w.Cm = ws.Cm; w.ldn = ldn; w.next_cols = (lookahead && next_panel) ? 1 : 0;
    dbg.begin(st);
    sb_w1_kernel<<<dim3(w.nblk, nz), 256, 0, st>>>(w);
    {
      const size_t w2smem = (size_t)(3 * NB2 + 2 * 64) * PP * sizeof(double);
      static PerDevice pd_w2cfg; size_t& w2cfg = pd_w2cfg.cur();
      if (!w2cfg) {
        APV_CUDA_TRY(cudaFuncSetAttribute(sb_w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w2smem));
        w2cfg = true;
      }
      sb_w2_kernel<<<dim3(w.nblk, nz), 1024, w2smem, st>>>(w);
    }
    dbg.end(st, 2);
    *launches += 2;
    tr.mark(st, 3);            // [3] look-ahead QR (the rank-2b update runs beside it on the side stream)
    // C22 -= V W^T + W V^T
    GemmArgs u{};
    u.batch = nz;
    u.A = ws.Z1 + (size_t)r * 2 * NB2; u.lda = 2 * NB2; u.strideA = (long long)n * 2 * NB2;
    u.B = ws.Z2 + (size_t)r * 2 * NB2; u.ldb = 2 * NB2; u.strideB = (long long)n * 2 * NB2;
    u.C = ws.Cm + (size_t)r * ldn + r; u.ldc = ldn; u.strideC = mstride;
    u.K = 2 * NB2; u.transB = 1; u.alpha = -1.0; u.beta = 1.0;
    u.bn = 64;                   // K = 64: two CTAs per SM hide the tile load / read-modify-write latency
    qr_done = false;
    if (lookahead && next_panel) {
      // (the next panel's columns -- its diagonal block and sub-diagonal panel -- were updated by sb_w2_kernel)
      // The panel QR goes first, in order on the main stream, so that its clusters find free SMs; the bulk of the
      // update follows on the side stream and fills the rest of the chip; the main stream joins it before Y.
      APV_CUDA_TRY(cudaEventRecord(ws.ev2[2], st));
      APV_TRY(launch_qr(r, st));
      qr_done = true;
      APV_CUDA_TRY(cudaStreamWaitEvent(ws.st2, ws.ev2[2], 0));
      u.A += (size_t)NB2 * 2 * NB2; u.B += (size_t)NB2 * 2 * NB2; u.C += (size_t)NB2 * ldn + NB2;
      u.M = npn - NB2; u.N = npn - NB2;
      u.tri = 1; u.mirror = 1;   // symmetric result: lower tiles computed, stored to both triangles
      APV_TRY(gemm_f64(u, ws.st2));
      APV_CUDA_TRY(cudaEventRecord(ws.ev2[3], ws.st2));
      tr.mark(st, 4);          // [4] waiting for the update after the look-ahead QR
      APV_CUDA_TRY(cudaStreamWaitEvent(st, ws.ev2[3], 0));
      ++*launches;
    } else {
      u.M = npn; u.N = npn;
      u.tri = 1; u.mirror = 1;
      dbg.begin(st);
      APV_TRY(gemm_f64(u, st));
      dbg.end(st, 3);
      ++*launches;
    }
  }
  tr.mark(st, 5);
  if (tr.on) {
    cudaStreamSynchronize(st);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i + 1 < tr.used; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tr.ev[i], tr.ev[i + 1]);
      acc[tr.cls[i]] += ms;
    }
    fprintf(stderr, "two-stage main-stream ms (look-ahead on): first QR %.2f | Y %.2f | W %.2f | look-ahead QR %.2f | wait for the update %.2f\n",
            acc[0], acc[1], acc[2], acc[3], acc[4]);
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev2[0], st));
  // ---- stage 2
  {
    const size_t tot = (size_t)n * LDB;
    sb_extract_band_kernel<<<dim3((unsigned)((tot + 255) / 256), nz), 256, 0, st>>>(ws.Cm, AB, n, ldn);
    ++*launches;
    if (n >= 3) {
      SbChase c;
      c.AB = AB; c.V2 = ws.Tm; c.gprog = prog; c.n = n; c.ldn = ldn; c.nz = nz; c.dbg = dbg.on ? dclk : nullptr;
      // groups in flight: a group lasts n / NB2 + 2 CGW time steps and the next one starts 2 CGW + 1 steps later
      const int ngroups = ceil_div(n - 2, CGW);
      // (measured at n = 4096: 20.6 ms from 18 CTAs per zone upwards, 20.7 with 16; every CTA more is an SM that the
      // statistics of the next block cannot use in the multi-block path)
      const int want = ceil_div(n / NB2 + 2 * CGW, 2 * CGW + 1);
      int G = std::max(1, std::min(std::min(sms / nz, want), ngroups));
      if (const char* e = getenv("APV_CHASE_G")) G = std::max(1, std::min(std::min(sms / nz, atoi(e)), ngroups));   // (experiment)
      const size_t smem = (size_t)CG_NSLOT * CG_PITCH * sizeof(double);
      static PerDevice pd_configured; size_t& configured = pd_configured.cur();
      if (!configured) {
        APV_CUDA_TRY(cudaFuncSetAttribute(sb2st_chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
      }
      void* args[] = {(void*)&c};
      dbg.begin(st);
      APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)sb2st_chase_kernel, dim3(G * nz), dim3(CG_THREADS), args, smem, st));
      dbg.end(st, 4);
      ++*launches;
    }
    sb_extract_tridiag_kernel<<<dim3(ceil_div(n, 256), nz), 256, 0, st>>>(AB, ws.dd, ws.ee, n);
    ++*launches;
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev2[1], st));
  APV_CUDA_TRY(cudaGetLastError());
  if (dbg.on) {
    fprintf(stderr, "two-stage ms: panel QR %.2f | Y = C22 V %.2f | W %.2f | rank-2b update %.2f | chase %.2f\n", dbg.acc[0],
            dbg.acc[1], dbg.acc[2], dbg.acc[3], dbg.acc[4]);
    cudaEventDestroy(dbg.e0);
    cudaEventDestroy(dbg.e1);
    long long hc[32];
    cudaMemcpy(hc, dclk, sizeof(hc), cudaMemcpyDeviceToHost);
    const double us = 1.0 / 1965.0;
    fprintf(stderr, "  panel QR us (rank 0): load %.0f | block reduce %.0f | deliver %.0f | cluster.sync %.0f | scalars %.0f | row loop %.0f | out %.0f\n",
            hc[0] * us, hc[1] * us, hc[2] * us, hc[3] * us, hc[4] * us, hc[5] * us, hc[6] * us);
    fprintf(stderr, "  chase us (first group): B-warp step %.0f | barrier %.0f ; storer: store+publish %.0f | barrier %.0f ; fetcher: wait %.0f | load %.0f | barrier %.0f\n",
            hc[8] * us, hc[11] * us, hc[12] * us, hc[15] * us, hc[17] * us, hc[18] * us, hc[19] * us);
#ifdef APV_CHASE_FINE
    fprintf(stderr, "  chase B warp us (first group): load B %.1f | y = B v %.1f | house %.1f | pair sync %.1f | B -= y v^T %.1f | column sums %.1f | B -= v u^T %.1f | store %.1f ; D warp: load %.1f | wait for the reflector %.1f | two-sided %.1f | store %.1f\n",
            hc[20] * us, hc[21] * us, hc[22] * us, hc[23] * us, hc[24] * us, hc[25] * us, hc[26] * us, hc[27] * us, hc[28] * us, hc[29] * us, hc[30] * us, hc[31] * us);
#endif
  }
  return OK;
}

// Eigenvectors of T (slot 4 of the inverse-iteration workspace) -> eigenvectors of the band matrix, in place.
int twostage_apply_q2(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n;
  if (n < 3) return OK;
  if (n > 32 * NB2 * Q2R) {
    snprintf(g_err, sizeof(g_err), "two-stage back-transformation: n = %d exceeds the built limit %d", n, 32 * NB2 * Q2R);
    return EINVAL_;
  }
  const size_t smem = (size_t)n * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (smem > 48 * 1024 && smem > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(sb2st_apply_q2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool dbg = getenv("APV_TS_DEBUG") != nullptr;
  if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
  if (ws.V <= Q2W_MAXV && !getenv("APV_Q2_PER_VECTOR")) {
    // wavefront of chains, lane = vector (the per-vector kernel below pays a CTA barrier per sweep; APV_Q2_PER_VECTOR=1)
    SbQ2 q;
    q.iv = ws.iv; q.V2 = ws.Tm; q.n = n; q.ldn = ws.ldn; q.Vp = ws.Vp; q.nz = ws.nz;
    q.ngrp = ceil_div(n - 2, NB2); q.nhalf = ws.Vp / 32;
    // the block counters live in the (then idle) Y slices of the band reduction
    q.prog = reinterpret_cast<int*>(ws.ts2 + (size_t)ws.nz * n * NB2);
    const int nprob = q.nz * q.nhalf;
    APV_CUDA_TRY(cudaMemsetAsync(q.prog, 0, (size_t)nprob * q.ngrp * sizeof(int), st));
    int dev = 0, sms = 0;
    APV_CUDA_TRY(cudaGetDevice(&dev));
    APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one warp per chain in flight: at most ngrp chains per problem, at most one CTA per SM
    const int warps = std::min(nprob * q.ngrp, sms * (Q2W_T / 32) / nprob * nprob);
    const int grid = std::max(1, ceil_div(std::max(warps, nprob), Q2W_T / 32));
    void* args[] = {(void*)&q};
    const size_t wsm = (size_t)(Q2W_T / 32) * 32 * 34 * sizeof(double);
    static PerDevice pd_wcfg; size_t& wcfg = pd_wcfg.cur();
    if (!wcfg) {
      APV_CUDA_TRY(cudaFuncSetAttribute(sb2st_apply_q2_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
      wcfg = 1;
    }
    APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)sb2st_apply_q2_wave_kernel, dim3(grid), dim3(Q2W_T), args, wsm, st));
  } else {
    sb2st_apply_q2_kernel<<<dim3(ws.V, ws.nz), Q2T, smem, st>>>(ws.iv, ws.Tm, n, ws.ldn, ws.Vp);
  }
  APV_CUDA_TRY(cudaGetLastError());
  if (dbg) {
    float ms = 0.f;
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "two-stage ms: apply Q2 %.2f\n", ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  ++*launches;
  return OK;
}

// Eigenvectors of the band matrix (inverse-iteration workspace) -> eigenvectors of C (rows of Zt).
int twostage_apply_q1(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, nz = ws.nz;
  int npanels = 0;
  for (int j0 = 0; n - j0 - NB2 >= 2; j0 += NB2) ++npanels;
  const int G = ceil_div(n, Q1RS);
  SbQ1 q;
  q.iv = ws.iv; q.VH = ws.VH; q.Zt = ws.Zt;
  q.Tall = ws.ts2 + (size_t)nz * n * NB2 * (2 + twostage_nsplit_max()) + (size_t)nz * ceil_div(n, 64) * NB2 * NB2;
  // partial S buffers and the barrier counters live in the (then idle) Y slices of the band reduction
  q.Sp = ws.ts2 + (size_t)nz * n * NB2;
  q.bar = reinterpret_cast<unsigned long long*>(q.Sp + (size_t)nz * 2 * G * NB2 * Q1VC);
  q.n = n; q.ldn = ws.ldn; q.V = ws.V; q.Vp = ws.Vp; q.nz = nz; q.npanels = npanels;
  const size_t smem = (size_t)(Q1RS * Q1VC + Q1RS * Q1VP + NB2 * Q1VC) * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (!configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(sb_apply_q1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  APV_CUDA_TRY(cudaMemsetAsync(q.bar, 0, nz * sizeof(unsigned long long), st));
  for (int v0 = 0; v0 < ws.V; v0 += Q1VC) {
    q.v0 = v0;
    void* args[] = {(void*)&q};
    APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)sb_apply_q1_kernel, dim3(G * nz), dim3(Q1T), args, smem, st));
    ++*launches;
  }
  return OK;
}

// Aggregation of the compact-WY factors of `gs` consecutive panels (gs <= 4) into ONE block reflector of width 32 gs:
//   H_1 ... H_gs = I - V T V^T,  V = [V_1 ... V_gs],  T upper triangular with the panels' own T_j on the diagonal and
//   T[0:32j, j] = -T[0:32j, 0:32j] (V_{1..j}^T V_j) T_j   (the dlarft recurrence on blocks).
// G = V^T V arrives as `nsplit` K-slice partials of a split GEMM.   grid (nz), 256 threads.
__global__ void __launch_bounds__(256) q1_aggregate_kernel(const double* __restrict__ Gpart, int nsplit, int split_cap,
                                                           const double* __restrict__ Tpan, int nz, int gs,
                                                           double* __restrict__ Tg, long long tg_stride) {
  extern __shared__ double agsm[];
  const int z = blockIdx.x, W = NB2 * gs, tid = threadIdx.x;
  double* G = agsm;                 // [W][W]   Gram matrix (sum of the K-slice partials)
  double* tmp = G + W * W;          // [W][32]
  double* T = Tg + (size_t)z * tg_stride;      // [W][128] in global memory (64 KB, L1/L2 resident)
  const double* gp = Gpart + (size_t)z * split_cap * 128 * 128;     // (the zone stride of the split GEMM's output)
  for (int e = tid; e < W * W; e += 256) {
    const int i = e / W, j = e % W;
    double acc = 0.0;
    for (int q = 0; q < nsplit; ++q) acc += gp[(size_t)q * 128 * 128 + i * 128 + j];
    G[e] = acc;
    T[i * 128 + j] = 0.0;
  }
  __syncthreads();
  for (int j = 0; j < gs; ++j) {          // the panels' own factors on the diagonal
    const double* tp = Tpan + ((size_t)j * nz + z) * NB2 * NB2;
    for (int e = tid; e < NB2 * NB2; e += 256) T[(j * NB2 + e / NB2) * 128 + j * NB2 + e % NB2] = tp[e];
  }
  __syncthreads();
  for (int j = 1; j < gs; ++j) {
    const int R = NB2 * j;               // rows above the diagonal block of block column j
    // tmp = G[0:R, j] T_j
    for (int e = tid; e < R * NB2; e += 256) {
      const int i = e / NB2, c = e % NB2;
      double acc = 0.0;
      for (int k = 0; k < NB2; ++k) acc = fma(G[i * W + j * NB2 + k], T[(j * NB2 + k) * 128 + j * NB2 + c], acc);
      tmp[e] = acc;
    }
    __syncthreads();
    // T[0:R, j] = -T[0:R, 0:R] tmp
    for (int e = tid; e < R * NB2; e += 256) {
      const int i = e / NB2, c = e % NB2;
      double acc = 0.0;
      for (int k = i; k < R; ++k) acc = fma(T[i * 128 + k], tmp[k * NB2 + c], acc);      // (T is upper triangular)
      T[i * 128 + j * NB2 + c] = -acc;
    }
    __syncthreads();
  }
}

// The same back-transformation for MANY vectors (full-spectrum requests, V > n / 8: BASELINE cfg-4), as tensor-core
// GEMMs on the n x Vp matrix of vectors X (slot 4 of the inverse-iteration workspace, row-major).  Four consecutive
// panels are aggregated into one block reflector of width 128 (q1_aggregate_kernel), so that every product has a
// full 128-row tile; per group, last to first:  S = V_g^T X[r:, :]  (128 x Vp, K = n - r),  U = T_g S,
// X[r:, :] -= V_g U  (rank-128 update).  The slab kernel above streams the reflectors once per chunk of 64 vectors (64
// launches of ~3 ms at V = n = 4096); here they are streamed once and the O(n^2 V) work runs on the DMMA pipe.
// X is transformed in place; the caller moves it on.
int twostage_apply_q1_gemm(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, nz = ws.nz, ldn = ws.ldn, Vp = ws.Vp;
  constexpr int GS = 4, GW = GS * NB2, NSPLIT = 16;
  int npanels = 0;
  for (int j0 = 0; n - j0 - NB2 >= 2; j0 += NB2) ++npanels;
  if (npanels == 0) return OK;
  const int ngroups = ceil_div(npanels, GS);
  const double* Tall = ws.ts2 + (size_t)nz * n * NB2 * (2 + twostage_nsplit_max()) + (size_t)nz * ceil_div(n, 64) * NB2 * NB2;
  const size_t need = (size_t)nz * ngroups * GW * GW + (size_t)nz * NSPLIT * GW * GW + 2 * (size_t)nz * GW * Vp;
  if (ws.q1agg_count < need) {
    if (ws.q1agg) cudaFree(ws.q1agg);
    ws.q1agg = nullptr; ws.q1agg_count = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&ws.q1agg, need * sizeof(double)));
    ws.q1agg_count = need;
  }
  double* Tg = ws.q1agg;                                        // [nz][ngroups][128][128]
  double* Gpart = Tg + (size_t)nz * ngroups * GW * GW;          // [nz][NSPLIT][128][128]
  double* S = Gpart + (size_t)nz * NSPLIT * GW * GW;            // [nz][128][Vp]
  double* U = S + (size_t)nz * GW * Vp;
  const long long tgs = (long long)ngroups * GW * GW, ss = (long long)GW * Vp;
  double* X = ws.iv + 4 * (size_t)n * Vp;
  const long long xs = 6LL * n * Vp;             // zone stride of the workspace
  const size_t agsm = (size_t)(GW * GW + GW * NB2) * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (!configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(q1_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agsm));
    configured = 1;
  }
  // ---- the aggregated T factors
  for (int g = 0; g < ngroups; ++g) {
    const int p0 = g * GS, gs = std::min(GS, npanels - p0), j0 = p0 * NB2, r = j0 + NB2, rows = n - r, Mg = NB2 * gs;
    const int kslice = round_up(ceil_div(rows, NSPLIT), 16);
    const int nsplit = ceil_div(rows, kslice);
    GemmArgs gr{};
    gr.batch = nz; gr.split = nsplit; gr.split_ktot = rows;
    gr.A = ws.VH + (size_t)j0 * ldn + r; gr.lda = ldn; gr.strideA = (long long)n * ldn; gr.splitA = kslice;
    gr.B = gr.A; gr.ldb = ldn; gr.strideB = gr.strideA; gr.splitB = kslice; gr.transB = 1;
    gr.C = Gpart; gr.ldc = GW; gr.strideC = (long long)NSPLIT * GW * GW; gr.splitC = (long long)GW * GW;
    gr.M = Mg; gr.N = Mg; gr.K = kslice; gr.alpha = 1.0; gr.beta = 0.0;
    APV_TRY(gemm_f64(gr, st));
    q1_aggregate_kernel<<<nz, 256, agsm, st>>>(Gpart, nsplit, NSPLIT, Tall + (size_t)p0 * nz * NB2 * NB2, nz, gs,
                                               Tg + (size_t)g * GW * GW, tgs);
    *launches += 2;
  }
  // ---- apply the block reflectors, last group first
  for (int g = ngroups - 1; g >= 0; --g) {
    const int p0 = g * GS, gs = std::min(GS, npanels - p0), j0 = p0 * NB2, r = j0 + NB2, rows = n - r, Mg = NB2 * gs;
    GemmArgs a{};
    a.batch = nz;
    a.A = ws.VH + (size_t)j0 * ldn + r; a.lda = ldn; a.strideA = (long long)n * ldn;      // V_g^T: Mg x rows
    a.B = X + (size_t)r * Vp; a.ldb = Vp; a.strideB = xs;
    a.C = S; a.ldc = Vp; a.strideC = ss;
    a.M = Mg; a.N = Vp; a.K = rows; a.alpha = 1.0; a.beta = 0.0; a.bn = 64;
    APV_TRY(gemm_f64(a, st));
    GemmArgs t{};
    t.batch = nz;
    t.A = Tg + (size_t)g * GW * GW; t.lda = GW; t.strideA = tgs;
    t.B = S; t.ldb = Vp; t.strideB = ss;
    t.C = U; t.ldc = Vp; t.strideC = ss;
    t.M = Mg; t.N = Vp; t.K = Mg; t.alpha = 1.0; t.beta = 0.0;
    APV_TRY(gemm_f64(t, st));
    GemmArgs u{};
    u.batch = nz;
    u.A = a.A; u.lda = ldn; u.strideA = a.strideA; u.transA = 1;                           // V_g: rows x Mg, stored Mg x rows
    u.B = U; u.ldb = Vp; u.strideB = ss;
    u.C = X + (size_t)r * Vp; u.ldc = Vp; u.strideC = xs;
    u.M = rows; u.N = Vp; u.K = Mg; u.alpha = -1.0; u.beta = 1.0;
    APV_TRY(gemm_f64(u, st));
    *launches += 3;
  }
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

}  // namespace apv
