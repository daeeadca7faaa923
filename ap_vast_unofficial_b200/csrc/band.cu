// syevd_f64, two-stage tridiagonalisation (eig_mode 3):  C = Q1 Cb Q1^T = Q1 Q2 T Q2^T Q1^T.
// The reference reaches the eigen-decomposition of C = L^-1 R_B L^-T through LAPACK dgees (Python/apvast.py:30);
// the one-stage reduction of tridiag.cu has to stream the trailing matrix once per column (BLAS-2, HBM-bound).
// Here the O(n^3) work is all BLAS-3 on the FP64 tensor cores and the BLAS-2 part runs on a band that stays in L2:
//
//   stage 1 (dense -> band, half bandwidth NB2 = 32), per panel of NB2 columns
//     sb_panel_qr_kernel   Householder QR of the sub-diagonal panel  C[r:, j0:j0+NB2]  (r = j0 + NB2) held in the
//                          DISTRIBUTED shared memory of one thread-block cluster per zone: one DSMEM exchange and one
//                          cluster barrier per column (a Gram row of the column gives norm, reflector, the row
//                          v^T P and the column of V^T V needed by the compact-WY factor T in the same reduction)
//     gemm (split-K)       Y = C22 V                                       (DMMA, skinny tile)
//     sb_w1/w2_kernel      X = Y T,  W = X - 1/2 V T^T (V^T X),  panels Z1 = [V | W], Z2 = [W | V]
//     gemm                 C22 -= Z1 Z2^T   (= V W^T + W V^T)              (DMMA)
//   stage 2 (band -> tridiagonal), sb2st_chase_kernel: bulge chasing, one WARP per sweep, the 32x32 blocks in
//     registers, sweeps pipelined two steps apart through progress counters in global memory (a sweep may run step
//     k once its predecessor has finished step k + 1); the band (n x 64 doubles per zone) lives in L2.
//   back-transformation of the V wanted eigenvectors: sb2st_apply_q2_kernel (stage-2 reflectors, one CTA per
//     vector, vector in shared memory), then the compact-WY kernel of jdiag.cu with the stage-1 reflectors.
#include <cooperative_groups.h>
#include <math.h>

#include <algorithm>

#include "engine.cuh"

namespace cg = cooperative_groups;

namespace apv {

namespace {

constexpr int NB2 = 32;          // half bandwidth of the intermediate band matrix == warp size (one lane per row)
constexpr int LDB = 2 * NB2;     // band storage: AB[j][d] = A[j + d][j], d = 0 .. 2 NB2 - 1 (band + bulge)
constexpr int QRT = 256;         // threads of the panel QR CTAs (8 warps, lane = panel column)
constexpr int QRW = QRT / 32;
constexpr int QR_MAXCS = 16;     // largest cluster
constexpr int PP = NB2 + 1;      // shared-memory pitch of the panel slab (odd: rows and columns conflict-free)
constexpr int CHT = 128;         // threads of the chase CTAs (4 sweeps per CTA)
constexpr int PROG_DONE = 0x7fffffff;

// ------------------------------------------------------------------------------------------------
// Stage 1, panel factorisation.  grid (CS, nz), cluster (CS, 1, 1).
struct SbPanel {
  double* Cm; double* VH; double* VP; double* tau; double* Tp;
  int n, ldn, j0, rows_per;
};

__global__ void __launch_bounds__(QRT, 1) sb_panel_qr_kernel(SbPanel a) {
  extern __shared__ __align__(16) double slab[];          // [rows_per][PP]
  __shared__ double xpart[2][QR_MAXCS][NB2];              // per-rank partial Gram rows (written by every rank)
  __shared__ double xrow[2][NB2];                         // the diagonal row of the current column (from its owner)
  __shared__ double red[QRW][NB2];
  __shared__ double Gs[NB2][PP];                          // V^T V above the diagonal (rank 0)
  __shared__ double taus[NB2];
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int z = blockIdx.y, n = a.n, ldn = a.ldn, j0 = a.j0, r = j0 + NB2, npn = n - r;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Cm = a.Cm + (size_t)z * n * ldn;
  const int s0 = min(npn, rank * a.rows_per), s1 = min(npn, s0 + a.rows_per), nr = s1 - s0;   // slab rows [s0, s1)
  for (int i = warp; i < nr; i += QRW) slab[i * PP + lane] = Cm[(size_t)(r + s0 + i) * ldn + j0 + lane];
  cluster.sync();             // every CTA of the cluster runs before its shared memory is written remotely
  // partial Gram row of column 0 over the rows strictly below its diagonal
  double acc = 0.0;
  for (int i = warp; i < nr; i += QRW)
    if (s0 + i > 0) acc = fma(slab[i * PP + lane], slab[i * PP], acc);
  for (int j = 0; j < NB2; ++j) {
    const int par = j & 1;
    red[warp][lane] = acc;
    __syncthreads();
    for (int k = warp; k < CS; k += QRW) {      // warp k delivers this CTA's partial to rank k (k + QRW, ..)
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < QRW; ++w) t += red[w][lane];
      cluster.map_shared_rank(&xpart[par][rank][0], k)[lane] = t;
    }
    if (j >= s0 && j < s1 && warp == QRW - 1) {      // owner of the diagonal row broadcasts it
      const double v = slab[(j - s0) * PP + lane];
      for (int k = 0; k < CS; ++k) cluster.map_shared_rank(&xrow[par][0], k)[lane] = v;
    }
    cluster.sync();
    double g = 0.0;
    for (int k = 0; k < CS; ++k) g += xpart[par][k][lane];
    const double rowj = (j < npn) ? xrow[par][lane] : 0.0;      // (no diagonal row: the column is empty)
    const double alpha = __shfl_sync(0xffffffffu, rowj, j), sigma = __shfl_sync(0xffffffffu, g, j);
    double beta = alpha, tau = 0.0, scale = 0.0;
    if (sigma > 0.0 && j < npn - 1) {
      beta = -copysign(hypot(alpha, sqrt(sigma)), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    const double q = fma(scale, g, rowj);     // lane > j: (v^T P)[lane];  lane < j: (V^T V)[lane][j]
    if (rank == 0 && warp == 0) {
      if (lane < j) Gs[lane][j] = q;
      if (lane == j) taus[j] = tau;
    }
    // update of the rows below the diagonal fused with the partial Gram row of column j + 1
    acc = 0.0;
    const double tq = tau * q;
    for (int i = warp; i < nr; i += QRW) {
      const int gr = s0 + i;
      if (gr <= j) continue;
      double x = slab[i * PP + lane];
      const double vr = slab[i * PP + j] * scale;
      if (lane > j) x = fma(-vr, tq, x);
      else if (lane == j) x = vr;
      if (lane >= j) slab[i * PP + lane] = x;
      const double xb = __shfl_sync(0xffffffffu, x, (j + 1) & 31);
      if (gr > j + 1) acc = fma(x, xb, acc);
    }
    if (j >= s0 && j < s1 && warp == 0) {            // diagonal row: R
      double x = slab[(j - s0) * PP + lane];
      if (lane > j) x -= tq;
      else if (lane == j) x = beta;
      slab[(j - s0) * PP + lane] = x;
    }
    // (the next iteration's barriers order these writes before any other warp reads them)
  }
  __syncthreads();
  // outputs: R (and zeros) into the panel of C, explicit V (row-major panel and as rows of VH)
  double* VP = a.VP + (size_t)z * n * NB2;
  for (int i = warp; i < nr; i += QRW) {
    const int gr = s0 + i;
    const double x = slab[i * PP + lane];
    Cm[(size_t)(r + gr) * ldn + j0 + lane] = (gr <= lane) ? x : 0.0;
    VP[(size_t)(r + gr) * NB2 + lane] = (gr > lane) ? x : (gr == lane ? 1.0 : 0.0);
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
      // T (dlarft, forward / columnwise): lane = row of T;  T[r][i] = -tau_i sum_{k=r}^{i-1} T[r][k] G[k][i]
      double T[NB2];
      double* Tp = a.Tp + (size_t)z * NB2 * NB2;
#pragma unroll
      for (int i = 0; i < NB2; ++i) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k)
          if (k >= lane) t = fma(T[k], Gs[k][i], t);
        T[i] = (lane == i) ? taus[i] : (lane < i ? -taus[i] * t : 0.0);
        Tp[lane * NB2 + i] = T[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// X = (sum_s Ypart[s]) T  and the partial  V^T X  of a block of 64 rows.   grid (ceil(npn/64), nz), 256 threads.
struct SbW {
  const double* Ypart; const double* VP; const double* Tp;
  double* X; double* Spart; double* Z1; double* Z2;
  int n, r, npn, nsplit, nblk;
};

__global__ void __launch_bounds__(256) sb_w1_kernel(SbW a) {
  __shared__ double Ys[64][PP], Vs[64][PP], Ts[NB2][PP];
  const int z = blockIdx.y, blk = blockIdx.x, n = a.n;
  const int row = threadIdx.x >> 2, cq = (threadIdx.x & 3) * 8;
  const int gr = a.r + blk * 64 + row;                       // global row
  const bool ok = gr < n;
  for (int i = threadIdx.x; i < NB2 * NB2; i += 256) Ts[i / NB2][i % NB2] = a.Tp[(size_t)z * NB2 * NB2 + i];
  double y[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) y[c] = 0.0;
  if (ok)
    for (int s = 0; s < a.nsplit; ++s) {
      const double* yp = a.Ypart + ((size_t)(z * a.nsplit + s) * n + gr) * NB2 + cq;
#pragma unroll
      for (int c = 0; c < 8; ++c) y[c] += yp[c];
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
  for (int k = 0; k < NB2; ++k) {
    const double yk = Ys[row][k];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = fma(yk, Ts[k][cq + c], x[c]);     // T is upper triangular (zeros stored)
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
  for (int i = 0; i < 64; ++i) {
    const double v = Vs[i][p];
#pragma unroll
    for (int c = 0; c < 4; ++c) s4[c] = fma(v, Ys[i][c0 + c], s4[c]);
  }
  double* sp = a.Spart + ((size_t)(z * a.nblk + blk) * NB2 + p) * NB2 + c0;
#pragma unroll
  for (int c = 0; c < 4; ++c) sp[c] = s4[c];
}

// W = X - V (1/2 T^T S),  S = sum of the partials;  Z1 = [V | W], Z2 = [W | V].
__global__ void __launch_bounds__(256) sb_w2_kernel(SbW a) {
  __shared__ double Ss[NB2][PP], Ms[NB2][PP], Ts[NB2][PP], Vs[64][PP];
  const int z = blockIdx.y, blk = blockIdx.x, n = a.n;
  for (int i = threadIdx.x; i < NB2 * NB2; i += 256) {
    double s = 0.0;
    const double* sp = a.Spart + (size_t)z * a.nblk * NB2 * NB2 + i;
    for (int b = 0; b < a.nblk; ++b) s += sp[(size_t)b * NB2 * NB2];
    Ss[i / NB2][i % NB2] = s;
    Ts[i / NB2][i % NB2] = a.Tp[(size_t)z * NB2 * NB2 + i];
  }
  const int row = threadIdx.x >> 2, cq = (threadIdx.x & 3) * 8;
  const int gr = a.r + blk * 64 + row;
  const bool ok = gr < n;
#pragma unroll
  for (int c = 0; c < 8; ++c) Vs[row][cq + c] = ok ? a.VP[((size_t)z * n + gr) * NB2 + cq + c] : 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < NB2 * NB2; i += 256) {
    const int p = i / NB2, c = i % NB2;
    double m = 0.0;
    for (int k = 0; k <= p; ++k) m = fma(Ts[k][p], Ss[k][c], m);       // (T^T S)[p][c]
    Ms[p][c] = 0.5 * m;
  }
  __syncthreads();
  if (!ok) return;
  double w[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = a.X[((size_t)z * n + gr) * NB2 + cq + c];
  for (int p = 0; p < NB2; ++p) {
    const double v = Vs[row][p];
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = fma(-v, Ms[p][cq + c], w[c]);
  }
  double* z1 = a.Z1 + ((size_t)z * n + gr) * 2 * NB2;
  double* z2 = a.Z2 + ((size_t)z * n + gr) * 2 * NB2;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double v = Vs[row][cq + c];
    z1[cq + c] = v; z1[NB2 + cq + c] = w[c];
    z2[cq + c] = w[c]; z2[NB2 + cq + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Band storage for stage 2: AB[j][d] = C[j + d][j] for d <= NB2 (lower band), zero for the bulge region.
__global__ void sb_extract_band_kernel(const double* __restrict__ Cm, double* __restrict__ AB, int n, int ldn) {
  const int z = blockIdx.y;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * LDB) return;
  const int j = (int)(t / LDB), d = (int)(t % LDB);
  double v = 0.0;
  if (d <= NB2 && j + d < n) v = Cm[(size_t)z * n * ldn + (size_t)(j + d) * ldn + j];
  AB[(size_t)z * n * LDB + t] = v;
}

__global__ void sb_extract_tridiag_kernel(const double* __restrict__ AB, double* __restrict__ dd, double* __restrict__ ee,
                                          int n) {
  const int z = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  dd[(size_t)z * n + j] = AB[((size_t)z * n + j) * LDB];
  ee[(size_t)z * n + j] = (j + 1 < n) ? AB[((size_t)z * n + j) * LDB + 1] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Stage 2: bulge chasing.  Sweep s annihilates column s of the band below the sub-diagonal with a reflector on
// rows [s+1, s+1+NB2) and chases the bulge down the band in steps of NB2 rows; every step k >= 1 works on the
// off-diagonal block B = A[q0:q0+NB2, r0:r0+NB2] (q0 = r0 + NB2) and the diagonal block D = A[q0:, q0:]:
//     B <- B H_prev;  new reflector H' from B[:, 0];  B[:, 1:] <- H' B[:, 1:];  D <- H' D H'.
// One warp per sweep, lane = row of the blocks (32 + 32 doubles per lane in registers), vectors are broadcast
// through a per-warp shared line, column sums use a butterfly transpose-reduction.
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
    beta = -copysign(hypot(alpha, sqrt(ss)), alpha);
    tau = (beta - alpha) / beta;
    v = (lane == 0) ? 1.0 : x / (alpha - beta);
  }
}

// D <- H D H for the symmetric block whose row `lane` is in D[] (full rows); v, tau = reflector; line = 2 x 32 doubles.
__device__ __forceinline__ void two_sided32(double (&D)[32], double v, double tau, int lane, double* line) {
  bcast_store(line, v, lane);
  double p = 0.0;
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double2 vv = *reinterpret_cast<const double2*>(line + c);
    p = fma(D[c], vv.x, p);
    p = fma(D[c + 1], vv.y, p);
  }
  p *= tau;
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
  double* AB; double* V2; int* prog;
  int n, ldn, nz;
};

__device__ __forceinline__ void chase_load_diag(const double* __restrict__ AB, int n, int q0, int lane, double (&D)[32]) {
  const int row = q0 + lane;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const int col = q0 + c;
    double v = 0.0;
    if (row < n && col < n) {
      v = (c <= lane) ? __ldcg(AB + (size_t)col * LDB + (lane - c)) : __ldcg(AB + (size_t)row * LDB + (c - lane));
    }
    D[c] = v;
  }
}

__device__ __forceinline__ void chase_store_diag(double* __restrict__ AB, int n, int q0, int lane, const double (&D)[32]) {
  const int row = q0 + lane;
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (c <= lane && row < n) AB[(size_t)(q0 + c) * LDB + (lane - c)] = D[c];
}

__device__ __forceinline__ void chase_wait(const int* prog, int s, int need, int lane) {
  if (s > 0) {
    if (lane == 0) {
      const volatile int* p = prog + (s - 1);
      while (*p < need) {
      }
      __threadfence();
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void chase_post(int* prog, int s, int done, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile int*>(prog + s) = done;
}

__global__ void __launch_bounds__(CHT, 1) sb2st_chase_kernel(SbChase a) {
  __shared__ __align__(16) double lines[CHT / 32][64];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nz = a.nz, z = blockIdx.x % nz, n = a.n;
  const int gw = (blockIdx.x / nz) * (CHT / 32) + wib, G = (gridDim.x / nz) * (CHT / 32);
  double* line = lines[wib];
  double* AB = a.AB + (size_t)z * n * LDB;
  double* V2 = a.V2 + (size_t)z * n * a.ldn;
  int* prog = a.prog + (size_t)z * n;
  for (int s = gw; s <= n - 3; s += G) {
    const int nsteps = 1 + (n - s - 2) / NB2;
    // ---- step 0: reflector of column s, two-sided update of the first diagonal block
    chase_wait(prog, s, 2, lane);
    int r0 = s + 1;
    double D[32];
    chase_load_diag(AB, n, r0, lane, D);
    const double x = (r0 + lane < n) ? __ldcg(AB + (size_t)s * LDB + 1 + lane) : 0.0;
    double v, tau, beta;
    warp_house(x, lane, v, tau, beta);
    if (lane == 0) AB[(size_t)s * LDB + 1] = beta;
    two_sided32(D, v, tau, lane, line);
    chase_store_diag(AB, n, r0, lane, D);
    if (r0 + lane < n) V2[(size_t)s * a.ldn + r0 + lane] = (lane == 0) ? tau : v;
    chase_post(prog, s, nsteps == 1 ? PROG_DONE : 1, lane);
    // ---- steps k >= 1
    for (int k = 1; k < nsteps; ++k) {
      const int q0 = r0 + NB2;
      chase_wait(prog, s, k + 2, lane);
      double B[32];
      const int row = q0 + lane;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        B[c] = (row < n) ? __ldcg(AB + (size_t)(r0 + c) * LDB + (NB2 + lane - c)) : 0.0;
      chase_load_diag(AB, n, q0, lane, D);
      // B <- B H_prev
      bcast_store(line, v, lane);
      double y = 0.0;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const double2 vv = *reinterpret_cast<const double2*>(line + c);
        y = fma(B[c], vv.x, y);
        y = fma(B[c + 1], vv.y, y);
      }
      y *= tau;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const double2 vv = *reinterpret_cast<const double2*>(line + c);
        B[c] = fma(-y, vv.x, B[c]);
        B[c + 1] = fma(-y, vv.y, B[c + 1]);
      }
      // new reflector from the first column of B
      warp_house(B[0], lane, v, tau, beta);
      B[0] = (lane == 0) ? beta : 0.0;
      // B[:, 1:] <- H' B[:, 1:]
      {
        double vals[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) vals[c] = v * B[c];
        const double u = tau * colsum32(vals, lane);       // lane c: tau v^T B[:, c]
        bcast_store(line, u, lane);
#pragma unroll
        for (int c = 2; c < 32; c += 2) {
          const double2 uu = *reinterpret_cast<const double2*>(line + c);
          B[c] = fma(-v, uu.x, B[c]);
          B[c + 1] = fma(-v, uu.y, B[c + 1]);
        }
        B[1] = fma(-v, line[1], B[1]);
      }
      if (row < n) {
#pragma unroll
        for (int c = 0; c < 32; ++c) AB[(size_t)(r0 + c) * LDB + (NB2 + lane - c)] = B[c];
      }
      two_sided32(D, v, tau, lane, line);
      chase_store_diag(AB, n, q0, lane, D);
      if (row < n) V2[(size_t)s * a.ldn + row] = (lane == 0) ? tau : v;
      chase_post(prog, s, k + 1 == nsteps ? PROG_DONE : k + 1, lane);
      r0 = q0;
    }
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


}  // namespace

size_t twostage_scratch_bytes(int n, int nz, int nsplit_max) {
  size_t d = 0;
  d += (size_t)nz * n * NB2;                          // VP
  d += (size_t)nz * nsplit_max * n * NB2;             // Ypart
  d += (size_t)nz * n * NB2;                          // X
  d += (size_t)nz * ceil_div(n, 64) * NB2 * NB2;      // Spart
  d += (size_t)nz * NB2 * NB2;                        // Tp
  d += (size_t)nz * n * LDB;                          // AB
  return d * sizeof(double) + (size_t)nz * n * sizeof(int);
}

int twostage_nsplit_max() { return 8; }

int twostage_run(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, nz = ws.nz;
  const long long mstride = (long long)n * ldn;
  const int nsm = twostage_nsplit_max();
  double* VP = ws.ts2;
  double* Ypart = VP + (size_t)nz * n * NB2;
  double* X = Ypart + (size_t)nz * nsm * n * NB2;
  double* Spart = X + (size_t)nz * n * NB2;
  double* Tp = Spart + (size_t)nz * ceil_div(n, 64) * NB2 * NB2;
  double* AB = Tp + (size_t)nz * NB2 * NB2;
  int* prog = reinterpret_cast<int*>(AB + (size_t)nz * n * LDB);
  int dev = 0, sms = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  APV_CUDA_TRY(cudaMemsetAsync(ws.tau, 0, (size_t)nz * n * sizeof(double), st));
  APV_CUDA_TRY(cudaMemsetAsync(prog, 0, (size_t)nz * n * sizeof(int), st));

  // ---- stage 1
  for (int j0 = 0; n - j0 - NB2 >= 2; j0 += NB2) {
    const int r = j0 + NB2, npn = n - r;
    // cluster size: smallest of 1, 2, 4, 8, 16 whose slab fits in shared memory
    const int max_rows = (200 * 1024) / (PP * (int)sizeof(double));
    int CS = 1;
    while (CS < QR_MAXCS && ceil_div(npn, CS) > max_rows) CS *= 2;
    if (npn > 64) CS = std::max(CS, 8);                       // spread the rows anyway: the column loop is latency-bound
    if (ceil_div(npn, CS) > max_rows) {
      snprintf(g_err, sizeof(g_err), "two-stage tridiagonalisation: n = %d exceeds the panel capacity", n);
      return EINVAL_;
    }
    SbPanel p;
    p.Cm = ws.Cm; p.VH = ws.VH; p.VP = VP; p.tau = ws.tau; p.Tp = Tp;
    p.n = n; p.ldn = ldn; p.j0 = j0; p.rows_per = ceil_div(npn, CS);
    const size_t smem = (size_t)p.rows_per * PP * sizeof(double);
    static thread_local size_t configured = 0;
    if (smem > configured) {
      APV_CUDA_TRY(cudaFuncSetAttribute(sb_panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)std::max(smem, (size_t)(200 * 1024))));
      APV_CUDA_TRY(cudaFuncSetAttribute(sb_panel_qr_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      configured = std::max(smem, (size_t)(200 * 1024));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, nz); cfg.blockDim = dim3(QRT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    APV_CUDA_TRY(cudaLaunchKernelEx(&cfg, sb_panel_qr_kernel, p));
    ++*launches;
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
    APV_TRY(gemm_f64(y, st));
    ++*launches;
    SbW w;
    w.Ypart = Ypart; w.VP = VP; w.Tp = Tp; w.X = X; w.Spart = Spart; w.Z1 = ws.Z1; w.Z2 = ws.Z2;
    w.n = n; w.r = r; w.npn = npn; w.nsplit = nsplit; w.nblk = ceil_div(npn, 64);
    sb_w1_kernel<<<dim3(w.nblk, nz), 256, 0, st>>>(w);
    sb_w2_kernel<<<dim3(w.nblk, nz), 256, 0, st>>>(w);
    *launches += 2;
    GemmArgs u{};                // C22 -= V W^T + W V^T
    u.batch = nz;
    u.A = ws.Z1 + (size_t)r * 2 * NB2; u.lda = 2 * NB2; u.strideA = (long long)n * 2 * NB2;
    u.B = ws.Z2 + (size_t)r * 2 * NB2; u.ldb = 2 * NB2; u.strideB = (long long)n * 2 * NB2;
    u.C = ws.Cm + (size_t)r * ldn + r; u.ldc = ldn; u.strideC = mstride;
    u.M = npn; u.N = npn; u.K = 2 * NB2; u.transB = 1; u.alpha = -1.0; u.beta = 1.0;
    APV_TRY(gemm_f64(u, st));
    ++*launches;
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev2[0], st));
  // ---- stage 2
  {
    const size_t tot = (size_t)n * LDB;
    sb_extract_band_kernel<<<dim3((unsigned)((tot + 255) / 256), nz), 256, 0, st>>>(ws.Cm, AB, n, ldn);
    ++*launches;
    if (n >= 3) {
      SbChase c;
      c.AB = AB; c.V2 = ws.Tm; c.prog = prog; c.n = n; c.ldn = ldn; c.nz = nz;
      // enough warps for the n / (2 NB2) sweeps that can be in flight, at most one CTA per SM
      const int want = ceil_div(ceil_div(n, NB2) + 2, CHT / 32);
      const int G = std::max(1, std::min(sms / nz, want));
      void* args[] = {(void*)&c};
      APV_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)sb2st_chase_kernel, dim3(G * nz), dim3(CHT), args, 0, st));
      ++*launches;
    }
    sb_extract_tridiag_kernel<<<dim3(ceil_div(n, 256), nz), 256, 0, st>>>(AB, ws.dd, ws.ee, n);
    ++*launches;
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev2[1], st));
  APV_CUDA_TRY(cudaGetLastError());
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
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(sb2st_apply_q2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  sb2st_apply_q2_kernel<<<dim3(ws.V, ws.nz), Q2T, smem, st>>>(ws.iv, ws.Tm, n, ws.ldn, ws.Vp);
  APV_CUDA_TRY(cudaGetLastError());
  ++*launches;
  return OK;
}

}  // namespace apv
