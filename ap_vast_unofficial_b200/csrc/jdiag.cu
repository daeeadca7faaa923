// S5  jdiag -- joint diagonalisation of (R_B, R_D) (reference jdiag, Python/apvast.py:20-36;
// spec Matlab/ControlMethods/jdiag.m:103-116):
//     Bc = chol(R_D + reg I)             -> chol_f64   (blocked right-looking, DMMA trailing update)
//     C  = Bc^-1 R_B Bc^-T               -> trsm_f64   (blocked, inverses of the diagonal super-blocks + DMMA GEMM)
//     C  = Q Lambda Q^T                  -> syevd_f64  (tridiagonalisation: two stages for n >= 1024 (band.cu: DMMA
//                                           band reduction + bulge chasing), one blocked Householder stage below
//                                           (tridiag.cu), shared-memory Jacobi for n <= 48; then top-V eigenpairs
//                                           of T by multisection and inverse iteration, back-transformed by the
//                                           reflectors)
//     U  = Bc^-T Q, columns sorted by descending eigenvalue
// The reference calls a non-symmetric real Schur (LAPACK dgees) on a matrix that is symmetric to
// rounding; here C is symmetrised and a symmetric eigensolver is used.  Only the V leading pairs
// are formed, because calculate_filter_spectra (apvast.py:406-414) consumes U[:, :V] only.
// Both zone problems are batched in every launch (blockIdx.y / blockIdx.z = zone).
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "engine.cuh"

namespace apv {

namespace {

constexpr int NB = 64;    // Cholesky / TRSM block size (diagonal blocks factorised / inverted in one CTA)
constexpr int SB = 4 * NB; // super-block: width of the delayed rank-SB trailing updates
constexpr int NBT = 32;   // tridiagonalisation panel width (one V and one W column per lane)

struct Ptr2 {
  const double* p[2];
};

// ------------------------------------------------------------------------------------------------
__global__ void prep_kernel(Ptr2 bright, Ptr2 dark, int ld_in, double* __restrict__ Cm, double* __restrict__ Lm,
                            int n, int ldn, double reg, const double* __restrict__ regv) {
  const int z = blockIdx.z;
  if (regv) reg = regv[z];      // per-zone loading computed on the device (norm-relative, apvast.py:25-27)
  const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = ((size_t)z * n + i) * ldn + j;
  Cm[o] = bright.p[z][(size_t)i * ld_in + j];
  Lm[o] = dark.p[z][(size_t)i * ld_in + j] + (i == j ? reg : 0.0);
}

// ------------------------------------------------------------------------------------------------
// Cholesky of one diagonal block (<= NB x NB, NB = 64) and the inverse of its factor.  grid (nz), 128 threads.
// The block is split 2 x 2 into 32 x 32 tiles; every tile operation is warp-synchronous with lane = row and the
// row in registers, so the only CTA barriers are the eight phase boundaries (the first version eliminated column
// by column with two CTA barriers per column: 94 us per block):
//   L11 = chol(A11) | L21 = A21 L11^-T, X11 = L11^-1 | A22 -= L21 L21^T | L22 = chol(A22) |
//   X22 = L22^-1, T = L21 X11 | X21 = -X22 T.
// A non-positive (or NaN) pivot records its 1-based global index in info[z][0] (first one wins) and is replaced
// by 1 so the launch sequence can finish; the host raises LinAlgError (numpy.linalg.cholesky, apvast.py:22-24).
constexpr int CDP = NB + 1;      // shared-memory pitch

// The kernel is launched 64 times per joint diagonalisation (n = 4096), each time on a cold instruction cache: a fully
// unrolled version (13 000 SASS instructions, 200 KB) spent its 49 us streaming its own code at ~2 bytes per cycle
// (in-kernel clocks: the phase times moved around with the data loads, their sum did not).  The eliminations are
// therefore ROLLED loops over a rotating register window -- the current column is always a[0], the update of the
// remaining columns and the shift are one FMA,  a[i] <- a[i+1] - m col[j+1+i]  -- and both halves of the 2 x 2
// blocking run through the same code (loop over h): ~1000 instructions.

// a[i] <- a[i+1] - m p[i]  for i < NI  (p: multipliers of the columns j+1 .. j+NI)
template <int NI>
__device__ __forceinline__ void rot_update(double (&a)[32], double m, const double* p) {
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = fma(-m, p[i], a[i + 1]);
}

// lane = row of a 32 x 32 symmetric tile, a[i] = entry (lane, i) (entries i <= lane matter).  Writes the row of the
// Cholesky factor to Srow[0 .. 32) (zeros above the diagonal) and the reciprocals of the diagonal to invd.
// col: 64 doubles of scratch.  Returns the first bad pivot or -1.
__device__ __forceinline__ int warp_chol32(double (&a)[32], int lane, double* col, double* invd, double* Srow) {
  int bad = -1;
  col[32 + lane] = 0.0;                     // multipliers of the columns beyond the tile
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    double d = __shfl_sync(0xffffffffu, a[0], j);
    if (!(d > 0.0)) {                       // also catches NaN
      if (bad < 0) bad = j;
      d = 1.0;
    }
    const double inv = rsqrt(d);
    const double l = (lane == j) ? d * inv : a[0] * inv;
    Srow[j] = (lane >= j) ? l : 0.0;
    __syncwarp();                           // the previous column's multipliers have been read
    col[lane] = l;
    if (lane == j) invd[j] = inv;
    __syncwarp();
    // (rows lane < j are finished: what they compute from here on is never used)
    if (j < 16) rot_update<31>(a, l, col + j + 1);
    else rot_update<15>(a, l, col + j + 1);
    a[31] = 0.0;
  }
  return bad;
}

// Forward substitution with a lower-triangular 32 x 32 tile Lt (pitch CDP, reciprocal diagonal invd), right-looking:
//   x_j = a[0] invd[j];  a <- a[1:] - x_j Lt[j+1:, j].   x_j is written to out[j * ostride].
// lane = row of A21 for L21 = A21 L11^-T (a = that row), lane = column of the inverse for X = Lt^-1 (a = e_lane).
__device__ __forceinline__ void warp_trsolve32(double (&a)[32], const double* Lt, const double* invd, double* out,
                                               int ostride) {
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const double x = a[0] * invd[j];
    out[j * ostride] = x;
    const double* lc = Lt + j;              // column j of the tile
    // (rows beyond the tile are clamped to its last row: they only feed window entries that are never used)
    if (j < 16) {
#pragma unroll
      for (int i = 0; i < 31; ++i) a[i] = fma(-x, lc[min(j + 1 + i, 31) * CDP], a[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 15; ++i) a[i] = fma(-x, lc[min(j + 1 + i, 31) * CDP], a[i + 1]);
    }
    a[31] = 0.0;
  }
}

#ifdef APV_CD_DEBUG
#define CD_INIT() long long cd_t[16]; int cd_n = 0; cd_t[cd_n++] = clock64()
#define CD_TICK(i) cd_t[cd_n++] = clock64()
#define CD_PRINT() do { if (threadIdx.x == 0 && blockIdx.x == 0 && k0 == 1024) { printf("chol_diag cycles:"); for (int q = 1; q < cd_n; ++q) printf(" %lld", cd_t[q] - cd_t[q - 1]); printf("\n"); } } while (0)
#else
#define CD_INIT()
#define CD_TICK(i)
#define CD_PRINT()
#endif
__global__ void __launch_bounds__(128) chol_diag_kernel(double* __restrict__ Lm, double* __restrict__ Dinv, int n,
                                                        int ldn, int k0, int nbk, int nblk, int* __restrict__ info) {
  extern __shared__ double cd_sm[];
  CD_INIT();
  double (*S)[CDP] = reinterpret_cast<double (*)[CDP]>(cd_sm);               // A -> L (lower)
  double (*X)[CDP] = reinterpret_cast<double (*)[CDP]>(cd_sm + NB * CDP);    // L^-1 (lower)
  __shared__ double Tm_[32][33];         // L21 X11
  __shared__ double invd[NB];
  __shared__ double cols[64];
  const int z = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* A = Lm + (size_t)z * n * ldn + (size_t)k0 * ldn + k0;
  {
    // all 32 loads of a thread in flight together (one after the other they cost one L2 round trip each)
    double v[NB * NB / 128];
#pragma unroll
    for (int u = 0; u < NB * NB / 128; ++u) {
      const int i = tid + 128 * u, r = i / NB, c = i % NB;
      v[u] = (c <= r) ? ((r < nbk && c < nbk) ? A[(size_t)r * ldn + c] : (r == c ? 1.0 : 0.0)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < NB * NB / 128; ++u) {
      const int i = tid + 128 * u, r = i / NB, c = i % NB;
      S[r][c] = v[u];
      X[r][c] = 0.0;
    }
  }
  __syncthreads();
  CD_TICK(0);
  double a[32];
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int o = 32 * h;
    if (warp == 0) {
      // ---- L_hh = chol(A_hh)  (h = 1: A22 has taken its update)
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = S[o + lane][o + c];
      const int bad = warp_chol32(a, lane, cols, invd + o, &S[o + lane][o]);
      if (bad >= 0 && o + bad < nbk && lane == 0 && info[z * 4] == 0) info[z * 4] = k0 + o + bad + 1;
    } else if (h == 1 && (warp == 1 || warp == 3)) {
      // ---- T = L21 X11 beside the factorisation of A22: lane = row, 16 columns per warp
#pragma unroll
      for (int k = 0; k < 32; ++k) a[k] = S[32 + lane][k];
      const int c0 = (warp == 1) ? 0 : 16;
#pragma unroll 1
      for (int cc = 0; cc < 16; ++cc) {
        const int c = c0 + cc;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          t0 = fma(a[k], X[k][c], t0);
          t1 = fma(a[k + 1], X[k + 1][c], t1);
        }
        Tm_[lane][c] = t0 + t1;
      }
    }
    __syncthreads();
    CD_TICK(1);
    if (warp == 2 || (warp == 1 && h == 0)) {
      // ---- warp 2: X_hh = L_hh^-1 (lane = column);  warp 1 (h = 0): L21 = A21 L11^-T (lane = row of A21)
      const bool inverse = warp == 2;
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = inverse ? (c == lane ? 1.0 : 0.0) : S[32 + lane][c];
      warp_trsolve32(a, &S[o][o], invd + o, inverse ? &X[o][o + lane] : &S[32 + lane][0], inverse ? CDP : 1);
    }
    __syncthreads();
    CD_TICK(2);
    if (h == 0) {
      // ---- A22 -= L21 L21^T: lane = row, every warp eight columns
#pragma unroll
      for (int k = 0; k < 32; ++k) a[k] = S[32 + lane][k];
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        const int c = warp * 8 + cc;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          t0 = fma(a[k], S[32 + c][k], t0);
          t1 = fma(a[k + 1], S[32 + c][k + 1], t1);
        }
        if (c <= lane) S[32 + lane][32 + c] -= t0 + t1;
      }
      __syncthreads();
      CD_TICK(3);
    }
  }
  // ---- X21 = -X22 T: lane = row, eight columns per warp
  {
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = X[32 + lane][32 + k];
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      const int c = warp * 8 + cc;
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k += 2) {
        t0 = fma(a[k], Tm_[k][c], t0);
        t1 = fma(a[k + 1], Tm_[k + 1][c], t1);
      }
      X[32 + lane][c] = -(t0 + t1);
    }
  }
  __syncthreads();
  CD_TICK(4);
  double* Di = Dinv + ((size_t)z * nblk + k0 / NB) * NB * NB;
  for (int i = tid; i < NB * NB; i += 128) {
    const int r = i / NB, c = i % NB;
    const bool in = c <= r && r < nbk && c < nbk;
    if (r < nbk && c < nbk) A[(size_t)r * ldn + c] = (c <= r) ? S[r][c] : 0.0;
    Di[i] = in ? X[r][c] : 0.0;
  }
  CD_TICK(5);
  CD_PRINT();
}

// ------------------------------------------------------------------------------------------------
// Inverses of the SB x SB diagonal blocks of the Cholesky factor, assembled from the NB x NB diagonal-block inverses
// (Dinv) by the block recurrence  X_ii = Dinv_i,  X_ij = -Dinv_i sum_{k=j..i-1} L_ik X_kj  (i > j).  With them a
// triangular solve is ONE product per super-block plus ONE trailing update (the NB-block substitution inside a
// super-block needed eight small launches).  grid (n / SB, nz), 256 threads (4 x 4 outputs each), 96 KB smem.
__global__ void __launch_bounds__(256) sb_inv_kernel(const double* __restrict__ Lm, const double* __restrict__ Dinv,
                                                     double* __restrict__ SBinv, int n, int ldn, int nblk) {
  extern __shared__ double si_sm[];
  double (*As)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(si_sm);
  double (*Bs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(si_sm + NB * (NB + 1));
  double (*Cs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(si_sm + 2 * NB * (NB + 1));
  const int sb = blockIdx.x, z = blockIdx.y, tid = threadIdx.x;
  const int s0 = sb * SB;
  const double* L = Lm + (size_t)z * n * ldn;
  const double* Di = Dinv + (size_t)z * nblk * NB * NB;
  double* X = SBinv + ((size_t)z * gridDim.x + sb) * SB * SB;          // [SB][SB] row-major
  const int ty = tid >> 4, tx = tid & 15;                                // outputs rows 4 ty.., cols 4 tx..
  constexpr int Q = SB / NB;
  // zero above the diagonal blocks, diagonal blocks from Dinv
  for (int i = tid; i < SB * SB; i += 256) {
    const int r = i / SB, c = i % SB, bi = r / NB, bj = c / NB;
    double v = 0.0;
    if (bi == bj && sb * Q + bi < nblk) v = Di[((size_t)(sb * Q + bi) * NB + (r % NB)) * NB + (c % NB)];
    X[i] = v;
  }
  __syncthreads();
  auto load_tile = [&](double (*T)[NB + 1], const double* src, int ld, int rows_valid, int cols_valid) {
    for (int i = tid; i < NB * NB; i += 256) {
      const int r = i / NB, c = i % NB;
      T[r][c] = (r < rows_valid && c < cols_valid) ? src[(size_t)r * ld + c] : 0.0;
    }
  };
  for (int i = 1; i < Q; ++i) {
    if (s0 + i * NB >= n) break;
    for (int j = i - 1; j >= 0; --j) {
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int k = j; k < i; ++k) {            // acc += L_ik X_kj
        __syncthreads();
        load_tile(As, L + (size_t)(s0 + i * NB) * ldn + s0 + k * NB, ldn, min(NB, n - s0 - i * NB), NB);
        load_tile(Bs, X + (size_t)(k * NB) * SB + j * NB, SB, NB, NB);
        __syncthreads();
#pragma unroll 8
        for (int q = 0; q < NB; ++q) {
          double av[4], bv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) av[a] = As[4 * ty + a][q];
#pragma unroll
          for (int b = 0; b < 4; ++b) bv[b] = Bs[q][4 * tx + b];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Cs[4 * ty + a][4 * tx + b] = acc[a][b];
      load_tile(As, Di + (size_t)(sb * Q + i) * NB * NB, NB, NB, NB);      // Dinv_i
      __syncthreads();
      double out[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[a][b] = 0.0;
#pragma unroll 8
      for (int q = 0; q < NB; ++q) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = As[4 * ty + a][q];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = Cs[q][4 * tx + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) out[a][b] = fma(av[a], bv[b], out[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) X[(size_t)(i * NB + 4 * ty + a) * SB + j * NB + 4 * tx + b] = -out[a][b];
      __threadfence_block();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// out = in^T (per zone), 32x32 tiles.
__global__ void transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int n, int ldn) {
  __shared__ double t[32][33];
  const size_t zo = (size_t)blockIdx.z * n * ldn;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    t[r][threadIdx.x] = (i < n && j < n) ? in[zo + (size_t)i * ldn + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;
    if (i < n && j < n) out[zo + (size_t)i * ldn + j] = t[threadIdx.x][r];
  }
}

// out = (in + in^T) / 2 (per zone).
__global__ void symmetrize_kernel(const double* __restrict__ in, double* __restrict__ out, int n, int ldn) {
  __shared__ double t[32][33];
  const size_t zo = (size_t)blockIdx.z * n * ldn;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;     // transposed tile
    t[r][threadIdx.x] = (i < n && j < n) ? in[zo + (size_t)i * ldn + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    if (i < n && j < n) out[zo + (size_t)i * ldn + j] = 0.5 * (in[zo + (size_t)i * ldn + j] + t[threadIdx.x][r]);
  }
}

// Upper triangle <- transpose of the lower triangle, in place (per zone).  grid (nt, nt, nz), tiles with bx >= by.
__global__ void mirror_lower_kernel(double* __restrict__ C, int n, int ldn) {
  __shared__ double t[32][33];
  if (blockIdx.x < blockIdx.y) return;              // only the tiles on and above the diagonal are written
  const size_t zo = (size_t)blockIdx.z * n * ldn;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;     // the mirrored (lower) tile
    t[r][threadIdx.x] = (i < n && j < n) ? C[zo + (size_t)i * ldn + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    if (i < n && j < n && j > i) C[zo + (size_t)i * ldn + j] = t[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------------------------
// Top-V eigenvalues of the symmetric tridiagonal (dd, ee) by multisection: one CTA per eigenvalue, 256 Sturm counts
// per pass (7 passes from the Gershgorin interval to working precision instead of 53 bisections).  Count
// recurrence as LAPACK dlaebz, with the quotient e^2 / t formed by a reciprocal and a product (the recurrence is a
// chain of n dependent divisions: this halves its latency; the count is insensitive to the last ulp of t).
// TPE threads per eigenvalue: 256 (one CTA per eigenvalue, grid (V, nz)) when few eigenvalues are wanted, 32 (one warp
// per eigenvalue, grid (V / 8, nz)) for full-spectrum requests where throughput matters.  256 threads; smem 2n doubles.
constexpr int BIS_T = 256;

// 1 / t to about one ulp: the 20-bit seed of the special-function unit and two Newton steps (four dependent FMAs).  The
// correctly rounded __drcp_rn costs ~40 cycles more per step of the recurrence below, which is a chain of n dependent
// reciprocals per shift (ncu: eig_bisect_kernel 23 instructions and ~137 cycles per step, 2 warps per SM sub-partition);
// the count is insensitive to the last ulp of t.  |t| >= pivmin >= DBL_MIN, so the seed neither overflows nor sees a
// subnormal; a subnormal result (|t| > 4.5e307) is flushed to zero, which is the right limit of e^2 / t.
__device__ __forceinline__ double rcp_fast(double t) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
  double e = fma(-t, y, 1.0);
  y = fma(y, e, y);
  e = fma(-t, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, int n,
                                           double x, double pivmin) {
  double t = d[0] - x;
  if (fabs(t) < pivmin) t = -pivmin;
  int c = (t <= 0.0);
#pragma unroll 4
  for (int i = 1; i < n; ++i) {
    t = fma(-e2[i - 1], rcp_fast(t), d[i] - x);
    if (fabs(t) < pivmin) t = -pivmin;
    c += (t <= 0.0);
  }
  return c;
}

template <int TPE>
__global__ void __launch_bounds__(BIS_T) eig_bisect_kernel(const double* __restrict__ dd, const double* __restrict__ ee,
                                                           double* __restrict__ lam, double* __restrict__ tnorm, int n,
                                                           int V) {
  extern __shared__ double sm[];
  double* d = sm;
  double* e2 = sm + n;
  __shared__ double red[40];
  __shared__ double sh[4];
  __shared__ int first[BIS_T / 32];
  constexpr int EPC = BIS_T / TPE;                 // eigenvalues per CTA
  const int z = blockIdx.y, v = blockIdx.x * EPC + threadIdx.x / TPE, gt = threadIdx.x % TPE;
  const double* gd = dd + (size_t)z * n;
  const double* ge = ee + (size_t)z * n;
  double gl = DBL_MAX, gu = -DBL_MAX, emax = 0.0, onenrm = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double di = gd[i];
    const double el = i > 0 ? fabs(ge[i - 1]) : 0.0, er = i < n - 1 ? fabs(ge[i]) : 0.0;
    d[i] = di;
    e2[i] = i < n - 1 ? ge[i] * ge[i] : 0.0;
    gl = fmin(gl, di - el - er);
    gu = fmax(gu, di + el + er);
    emax = fmax(emax, er * er);
    onenrm = fmax(onenrm, fabs(di) + el + er);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    onenrm = fmax(onenrm, __shfl_xor_sync(0xffffffffu, onenrm, o));
  }
  if (lane == 0) { red[warp] = gl; red[8 + warp] = gu; red[16 + warp] = emax; red[24 + warp] = onenrm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BIS_T / 32; ++w) {
      red[0] = fmin(red[0], red[w]); red[8] = fmax(red[8], red[8 + w]);
      red[16] = fmax(red[16], red[16 + w]); red[24] = fmax(red[24], red[24 + w]);
    }
    const double tn = fmax(fabs(red[0]), fabs(red[8]));
    sh[0] = red[0] - 2.1 * tn * DBL_EPSILON * n - 4.2 * DBL_MIN;   // dstebz widening
    sh[1] = red[8] + 2.1 * tn * DBL_EPSILON * n + 4.2 * DBL_MIN;
    sh[2] = DBL_MIN * fmax(1.0, red[16]);                          // pivmin
    sh[3] = red[24];
    if (blockIdx.x == 0) { tnorm[z * 4 + 0] = red[24]; tnorm[z * 4 + 1] = sh[2]; tnorm[z * 4 + 2] = tn; }
  }
  __syncthreads();
  if (TPE == 32 && v >= V) return;                 // (TPE == BIS_T: the grid is exact, the CTA stays convergent)
  const int k = n - v;            // k-th smallest (1-based)
  double lo = sh[0], hi = sh[1];
  const double pivmin = sh[2];
  const double atol = DBL_EPSILON * fmax(fabs(lo), fabs(hi));
  for (int it = 0; it < 100; ++it) {
    const double width = hi - lo;
    if (width <= fmax(fmax(atol, pivmin), 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)))) break;
    const double h = width / (double)(TPE + 1);
    const double x = lo + (gt + 1) * h;
    const int c = sturm_count(d, e2, n, x, pivmin);
    const unsigned mask = __ballot_sync(0xffffffffu, c >= k);
    int f;
    if (TPE == 32) {
      f = mask ? __ffs(mask) - 1 : TPE;
    } else {
      if (lane == 0) first[warp] = mask ? warp * 32 + __ffs(mask) - 1 : TPE;
      __syncthreads();
      f = TPE;
#pragma unroll
      for (int w = 0; w < BIS_T / 32; ++w) f = min(f, first[w]);
      __syncthreads();
    }
    // the eigenvalue lies in (x_{f-1}, x_f] with x_{-1} = lo, x_{TPE} = hi
    const double nlo = f > 0 ? lo + f * h : lo;
    const double nhi = f < TPE ? lo + (f + 1) * h : hi;
    if (!(nhi - nlo < width)) break;     // no progress at working precision
    lo = nlo; hi = nhi;
  }
  if (gt == 0) lam[(size_t)z * V + v] = 0.5 * (lo + hi);
}

// Shifts for inverse iteration: ascending walk, close values pushed apart by 10 eps |x| (LAPACK dstein).
__global__ void eig_shift_kernel(const double* __restrict__ lam, double* __restrict__ shift, int V) {
  const int z = blockIdx.x;
  if (threadIdx.x != 0) return;
  double xjm = 0.0;
  for (int v = V - 1; v >= 0; --v) {        // lam is descending -> walk ascending
    double xj = lam[(size_t)z * V + v];
    if (v < V - 1) {
      const double pertol = 10.0 * fabs(DBL_EPSILON * xj);
      if (xj - xjm < pertol) xj = xjm + pertol;
    }
    shift[(size_t)z * V + v] = xj;
    xjm = xj;
  }
}

// ------------------------------------------------------------------------------------------------
// Inverse iteration, one thread per eigenvector (LAPACK dstein with dlagtf / dlagts(job=-1) restated).
// Work arrays are interleaved over vectors: arr[i * Vp + v].  grid (Vp/32, nz), 32 threads.
__device__ __forceinline__ double urand(unsigned long long& s) {
  s += 0x9E3779B97F4A7C15ull;
  unsigned long long x = s;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void __launch_bounds__(32) eig_invit_kernel(const double* __restrict__ dd, const double* __restrict__ ee,
                                                       const double* __restrict__ shift,
                                                       const double* __restrict__ tnorm, double* __restrict__ iv,
                                                       int* __restrict__ info, int n, int V, int Vp) {
  const int z = blockIdx.y;
  const int v = blockIdx.x * 32 + threadIdx.x;
  if (v >= V) return;
  const double* d = dd + (size_t)z * n;
  const double* e = ee + (size_t)z * n;
  double* base = iv + (size_t)z * 6 * n * Vp;
  double* A = base + 0 * (size_t)n * Vp + v;
  double* B = base + 1 * (size_t)n * Vp + v;
  double* C = base + 2 * (size_t)n * Vp + v;
  double* D2 = base + 3 * (size_t)n * Vp + v;
  double* X = base + 4 * (size_t)n * Vp + v;
  double* IN = base + 5 * (size_t)n * Vp + v;
  const size_t S = (size_t)Vp;
  const double lamv = shift[(size_t)z * V + v];
  const double onenrm = tnorm[z * 4 + 0];
  const double eps = DBL_EPSILON;
  unsigned long long seed = 0x1234567ull + 7919ull * (unsigned long long)v + 104729ull * (unsigned long long)z;

  if (n == 1) {
    X[0] = 1.0;
    return;
  }
  // ---- dlagtf: T - lam I = P L U
  double ak = d[0] - lamv;
  double bk = e[0];
  double tolmax = 0.0;
  double scale1 = fabs(ak) + fabs(bk);
  for (int k = 0; k < n - 1; ++k) {
    double ak1 = d[k + 1] - lamv;
    const double ck = e[k];
    const double bk1 = (k < n - 2) ? e[k + 1] : 0.0;
    double scale2 = fabs(ck) + fabs(ak1);
    if (k < n - 2) scale2 += fabs(bk1);
    const double piv1 = (ak == 0.0) ? 0.0 : fabs(ak) / scale1;
    double cout, dout = 0.0, inflag = 0.0, bnext = bk1;
    if (ck == 0.0) {
      scale1 = scale2;
      cout = 0.0;
    } else {
      const double piv2 = fabs(ck) / scale2;
      if (piv2 <= piv1) {
        scale1 = scale2;
        cout = ck / ak;
        ak1 -= cout * bk;
      } else {
        inflag = 1.0;
        const double mult = ak / ck;
        ak = ck;
        const double temp = ak1;
        ak1 = bk - mult * temp;
        if (k < n - 2) {
          dout = bk1;
          bnext = -mult * dout;
        }
        bk = temp;
        cout = mult;
      }
    }
    A[k * S] = ak; B[k * S] = bk; C[k * S] = cout; D2[k * S] = dout; IN[k * S] = inflag;
    tolmax = fmax(tolmax, fmax(fabs(ak), fmax(fabs(bk), fabs(dout))));
    ak = ak1;
    bk = bnext;
  }
  A[(size_t)(n - 1) * S] = ak;
  tolmax = fmax(tolmax, fabs(ak));
  const double alast = ak;
  double tol = tolmax * eps;
  if (tol == 0.0) tol = eps;
  const double sfmin = DBL_MIN, bignum = 1.0 / DBL_MIN;

  // ---- iterations
  for (int i = 0; i < n; ++i) X[i * S] = urand(seed);
  const double dtpcrt = sqrt(0.1 / n);
  int nrmchk = 0, its = 0;
  bool ok = false;
  while (its < 8) {
    ++its;
    double xmax = 0.0;
    for (int i = 0; i < n; ++i) xmax = fmax(xmax, fabs(X[i * S]));
    const double scl = n * onenrm * fmax(eps, fabs(alast)) / xmax;
    // forward elimination (with the scaling folded in)
    double prev = X[0] * scl;
    for (int k = 1; k < n; ++k) {
      const double xk = X[k * S] * scl;
      const double c = C[(k - 1) * S];
      if (IN[(k - 1) * S] == 0.0) {
        X[(k - 1) * S] = prev;
        prev = xk - c * prev;
      } else {
        X[(k - 1) * S] = xk;
        prev = prev - c * xk;
      }
    }
    X[(size_t)(n - 1) * S] = prev;
    // back substitution with pivot perturbation (dlagts job = -1)
    double y1 = 0.0, y2 = 0.0;        // x[k+1], x[k+2]
    for (int k = n - 1; k >= 0; --k) {
      double temp = X[k * S];
      if (k <= n - 3) temp = temp - B[k * S] * y1 - D2[k * S] * y2;
      else if (k == n - 2) temp = temp - B[k * S] * y1;
      double akk = A[k * S];
      double pert = copysign(tol, akk);
      for (;;) {
        const double absak = fabs(akk);
        if (absak < 1.0) {
          if (absak < sfmin) {
            if (absak == 0.0 || fabs(temp) * sfmin > absak) { akk += pert; pert *= 2.0; continue; }
            temp *= bignum; akk *= bignum;
          } else if (fabs(temp) > absak * bignum) { akk += pert; pert *= 2.0; continue; }
        }
        break;
      }
      const double xk = temp / akk;
      X[k * S] = xk;
      y2 = y1; y1 = xk;
    }
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) nrm = fmax(nrm, fabs(X[i * S]));
    if (nrm < dtpcrt) continue;
    if (++nrmchk < 3) continue;
    ok = true;
    break;
  }
  if (!ok) atomicOr(&info[z * 4 + 1], 1);
  double ss = 0.0;
  for (int i = 0; i < n; ++i) { const double t = X[i * S]; ss += t * t; }
  const double inv = 1.0 / sqrt(ss);
  for (int i = 0; i < n; ++i) X[i * S] *= inv;
}

// Same algorithm with the per-vector work arrays in shared memory: one CTA per eigenvector, thread 0 runs the
// sequential recurrences at shared-memory latency, the other threads do the fills, norms and scalings.
// Used when 6 n doubles + n flags fit in one SM's shared memory (n <= ~4600).   grid (V, nz), 128 threads.
__device__ __forceinline__ double hash_uniform(unsigned long long key) {
  unsigned long long x = key + 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__device__ __forceinline__ double block_max128(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
}

__global__ void __launch_bounds__(128) eig_invit_smem_kernel(const double* __restrict__ dd, const double* __restrict__ ee,
                                                             const double* __restrict__ shift,
                                                             const double* __restrict__ tnorm, double* __restrict__ iv,
                                                             int* __restrict__ info, int n, int V, int Vp) {
  extern __shared__ __align__(16) double ism[];
  double* A = ism;
  double* B = A + n;
  double* C = B + n;
  double* D2 = C + n;
  double* X = D2 + n;
  double* RA = X + n;          // reciprocals of the pivots
  unsigned char* IN = reinterpret_cast<unsigned char*>(RA + n);
  __shared__ double red[8];
  __shared__ double sh[4];
  const int v = blockIdx.x, z = blockIdx.y;
  const double* d = dd + (size_t)z * n;
  const double* e = ee + (size_t)z * n;
  double* Xout = iv + (size_t)z * 6 * n * Vp + 4 * (size_t)n * Vp + v;
  const double lamv = shift[(size_t)z * V + v];
  const double onenrm = tnorm[z * 4 + 0];
  const double eps = DBL_EPSILON;
  if (n == 1) {
    if (threadIdx.x == 0) Xout[0] = 1.0;
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    A[i] = d[i] - lamv;
    B[i] = (i < n - 1) ? e[i] : 0.0;
    C[i] = (i < n - 1) ? e[i] : 0.0;
    D2[i] = 0.0;
    IN[i] = 0;
    X[i] = hash_uniform(((unsigned long long)(z * 131071 + v) << 32) ^ (unsigned long long)i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {      // dlagtf; the running A[k], B[k] are carried in registers, one division per step
    double tolmax = 0.0;
    double ak = A[0], bk = B[0];
    double scale1 = fabs(ak) + fabs(bk);
    for (int k = 0; k < n - 1; ++k) {
      const double ck = C[k], ak1 = A[k + 1];
      const double bk1 = (k < n - 2) ? B[k + 1] : 0.0;
      const double scale2 = fabs(ck) + fabs(ak1) + fabs(bk1);
      double na = ak1, nb = bk1, oa = ak, ob = bk, d2k = 0.0;
      if (ck == 0.0) {
        scale1 = scale2;
      } else if (ak != 0.0 && fabs(ck) * scale1 <= fabs(ak) * scale2) {      // piv2 <= piv1 without the quotients
        scale1 = scale2;
        const double m = ck / ak;
        C[k] = m;
        na = ak1 - m * bk;
      } else {
        IN[k] = 1;
        const double mult = ak / ck;
        oa = ck;
        na = bk - mult * ak1;
        if (k < n - 2) {
          d2k = bk1;
          nb = -mult * bk1;
          D2[k] = d2k;
        }
        ob = ak1;
        C[k] = mult;
        A[k] = oa;
        B[k] = ob;
      }
      tolmax = fmax(tolmax, fmax(fabs(oa), fmax(fabs(ob), fabs(d2k))));
      A[k + 1] = na;
      if (k < n - 2) B[k + 1] = nb;
      ak = na; bk = nb;
    }
    tolmax = fmax(tolmax, fabs(A[n - 1]));
    double tol = tolmax * eps;
    if (tol == 0.0) tol = eps;
    sh[0] = tol;
    sh[1] = A[n - 1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) RA[i] = 1.0 / A[i];      // (inf for a zero pivot: never used then)
  __syncthreads();
  const double tol = sh[0], alast = sh[1];
  const double sfmin = DBL_MIN, bignum = 1.0 / DBL_MIN;
  const double dtpcrt = sqrt(0.1 / n);
  int nrmchk = 0, its = 0;
  bool ok = false;
  while (its < 8) {
    ++its;
    double xm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) xm = fmax(xm, fabs(X[i]));
    xm = block_max128(xm, red);
    const double scl = n * onenrm * fmax(eps, fabs(alast)) / xm;
    for (int i = threadIdx.x; i < n; i += blockDim.x) X[i] *= scl;
    __syncthreads();
    if (threadIdx.x == 0) {
      // forward elimination (dlagts job = -1), the running entry carried in a register
      double xp = X[0];
      for (int k = 1; k < n; ++k) {
        double xk = X[k];
        const double c = C[k - 1];
        if (IN[k - 1] == 0) {
          xk = xk - c * xp;
        } else {
          const double t = xp;
          xp = xk;
          xk = t - c * xk;
        }
        X[k - 1] = xp;
        xp = xk;
      }
      X[n - 1] = xp;
      // back substitution with pivot perturbation; regular pivots are applied through their reciprocals
      double x1 = 0.0, x2 = 0.0;          // X[k + 1], X[k + 2]  (B[n - 1] = D2[n - 2] = D2[n - 1] = 0)
      for (int k = n - 1; k >= 0; --k) {
        double temp = X[k] - B[k] * x1 - D2[k] * x2;
        double akk = A[k];
        const double absk = fabs(akk);
        double xk;
        if (absk >= 1.0 || (absk >= sfmin && !(fabs(temp) > absk * bignum))) {
          xk = temp * RA[k];
        } else {
          double pert = copysign(tol, akk);
          for (;;) {
            const double absak = fabs(akk);
            if (absak < 1.0) {
              if (absak < sfmin) {
                if (absak == 0.0 || fabs(temp) * sfmin > absak) { akk += pert; pert *= 2.0; continue; }
                temp *= bignum; akk *= bignum;
              } else if (fabs(temp) > absak * bignum) { akk += pert; pert *= 2.0; continue; }
            }
            break;
          }
          xk = temp / akk;
        }
        X[k] = xk;
        x2 = x1; x1 = xk;
      }
    }
    __syncthreads();
    double nrm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) nrm = fmax(nrm, fabs(X[i]));
    nrm = block_max128(nrm, red);
    if (nrm < dtpcrt) continue;
    if (++nrmchk < 3) continue;
    ok = true;
    break;
  }
  if (!ok && threadIdx.x == 0) atomicOr(&info[z * 4 + 1], 1);
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fma(X[i], X[i], ss);
  ss = warp_sum(ss);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const double inv = 1.0 / sqrt(red[0] + red[1] + red[2] + red[3]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) Xout[(size_t)i * Vp] = X[i] * inv;
}

// Re-orthogonalise eigenvectors whose eigenvalues are (nearly) degenerate: modified Gram-Schmidt inside each
// run of consecutive eigenvalues closer than ctol * |T|.  One CTA per zone.
__global__ void __launch_bounds__(256) eig_cluster_mgs_kernel(const double* __restrict__ lam,
                                                              const double* __restrict__ tnorm,
                                                              double* __restrict__ iv, int n, int V, int Vp, double ctol) {
  __shared__ double red[40];
  const int z = blockIdx.x;
  double* X = iv + (size_t)z * 6 * n * Vp + 4 * (size_t)n * Vp;
  const double thr = ctol * tnorm[z * 4 + 0];
  int start = V - 1;                      // ascending walk: v = V-1 (smallest) .. 0
  for (int v = V - 2; v >= 0; --v) {
    const bool close = fabs(lam[(size_t)z * V + v] - lam[(size_t)z * V + v + 1]) <= thr;
    if (!close) { start = v; continue; }
    for (int u = start; u > v; --u) {     // orthogonalise x_v against x_u, u in the same cluster
      double s = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) s += X[(size_t)i * Vp + v] * X[(size_t)i * Vp + u];
      s = block_sum(s, red);
      for (int i = threadIdx.x; i < n; i += blockDim.x) X[(size_t)i * Vp + v] -= s * X[(size_t)i * Vp + u];
      __syncthreads();
    }
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double t = X[(size_t)i * Vp + v]; s += t * t; }
    s = block_sum(s, red);
    const double inv = 1.0 / sqrt(s);
    for (int i = threadIdx.x; i < n; i += blockDim.x) X[(size_t)i * Vp + v] *= inv;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Back-transformation q = Q z = H_0 H_1 ... H_{n-2} z with the reflectors grouped in blocks of WYB:
// H_{j0} ... H_{j0+WYB-1} = I - V T V^T (compact WY, LAPACK dlarft forward/columnwise).
// (1) wy_tfactor_kernel: one CTA per block forms the Gram matrix V^T V and T.   grid (ceil(n/WYB), nz)
constexpr int WYB = 8;
constexpr int BTT = 1024;   // threads of the back-transformation / back-solve CTAs

__global__ void __launch_bounds__(256) wy_tfactor_kernel(const double* __restrict__ VH, const double* __restrict__ tau,
                                                         double* __restrict__ Tf, int n, int ldn) {
  __shared__ double G[WYB][WYB];
  __shared__ double red[8][WYB * (WYB + 1) / 2];
  const int blk = blockIdx.x, z = blockIdx.y, j0 = blk * WYB;
  const int nb = min(WYB, n - j0);
  const double* vh = VH + (size_t)z * n * ldn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[WYB * (WYB + 1) / 2];
#pragma unroll
  for (int p = 0; p < WYB * (WYB + 1) / 2; ++p) acc[p] = 0.0;
  for (int i = j0 + 1 + threadIdx.x; i < n; i += blockDim.x) {
    double v[WYB];
#pragma unroll
    for (int c = 0; c < WYB; ++c) v[c] = (c < nb) ? vh[(size_t)(j0 + c) * ldn + i] : 0.0;
    int p = 0;
#pragma unroll
    for (int c = 0; c < WYB; ++c)
#pragma unroll
      for (int d = 0; d <= c; ++d) { acc[p] = fma(v[c], v[d], acc[p]); ++p; }
  }
#pragma unroll
  for (int p = 0; p < WYB * (WYB + 1) / 2; ++p) {
    const double t = warp_sum(acc[p]);
    if (lane == 0) red[warp][p] = t;
  }
  __syncthreads();
  if (threadIdx.x < WYB * (WYB + 1) / 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    int c = 0, p = threadIdx.x;
    while (p > c) { p -= c + 1; ++c; }
    G[c][p] = t;
    G[p][c] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double T[WYB][WYB];
    for (int c = 0; c < WYB; ++c)
      for (int d = 0; d < WYB; ++d) T[c][d] = 0.0;
    for (int i = 0; i < nb; ++i) {
      const double ti = tau[(size_t)z * n + j0 + i];
      T[i][i] = ti;
      // T(0:i, i) = -tau_i * T(0:i, 0:i) * (V(:, 0:i)^T v_i)
      for (int r = 0; r < i; ++r) {
        double sacc = 0.0;
        for (int k = r; k < i; ++k) sacc += T[r][k] * G[k][i];
        T[r][i] = -ti * sacc;
      }
    }
    double* out = Tf + ((size_t)z * gridDim.x + blk) * WYB * WYB;
    for (int c = 0; c < WYB; ++c)
      for (int d = 0; d < WYB; ++d) out[c * WYB + d] = T[c][d];
  }
}

// (2) one CTA per eigenvector: z <- (I - V T V^T) z for the blocks in descending order; one barrier per block.
__global__ void __launch_bounds__(BTT) eig_backtransform_kernel(const double* __restrict__ iv,
                                                                const double* __restrict__ VH,
                                                                const double* __restrict__ Tf,
                                                                double* __restrict__ Zt, int n, int ldn, int V, int Vp) {
  extern __shared__ double xs[];
  __shared__ double red[2][BTT / 32][WYB];
  const int v = blockIdx.x, z = blockIdx.y;
  const double* X = iv + (size_t)z * 6 * n * Vp + 4 * (size_t)n * Vp + v;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = X[(size_t)i * Vp];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* vh = VH + (size_t)z * n * ldn;
  const int nblk = (n + WYB - 1) / WYB;
  const double* tf = Tf + (size_t)z * nblk * WYB * WYB;
  int par = 0;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int j0 = blk * WYB;
    // fixed ownership i == threadIdx.x (mod BTT) so xs needs no barrier between blocks
    int i0 = threadIdx.x;
    if (i0 <= j0) i0 += ((j0 - i0) / BTT + 1) * BTT;
    double s[WYB];
#pragma unroll
    for (int c = 0; c < WYB; ++c) s[c] = 0.0;
#pragma unroll 4
    for (int i = i0; i < n; i += BTT) {
      const double xi = xs[i];
#pragma unroll
      for (int c = 0; c < WYB; ++c)
        if (j0 + c < n) s[c] = fma(vh[(size_t)(j0 + c) * ldn + i], xi, s[c]);
    }
#pragma unroll
    for (int c = 0; c < WYB; ++c) {
      const double t = warp_sum(s[c]);
      if (lane == 0) red[par][warp][c] = t;
    }
    __syncthreads();
    double sv[WYB], tv[WYB];
#pragma unroll
    for (int c = 0; c < WYB; ++c) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < BTT / 32; ++w) t += red[par][w][c];
      sv[c] = t;
    }
    const double* T = tf + (size_t)blk * WYB * WYB;
#pragma unroll
    for (int c = 0; c < WYB; ++c) {
      double t = 0.0;
#pragma unroll
      for (int d = 0; d < WYB; ++d)
        if (d >= c) t = fma(__ldg(T + c * WYB + d), sv[d], t);
      tv[c] = t;
    }
#pragma unroll 4
    for (int i = i0; i < n; i += BTT) {
      double xi = xs[i];
#pragma unroll
      for (int c = 0; c < WYB; ++c)
        if (j0 + c < n) xi = fma(-tv[c], vh[(size_t)(j0 + c) * ldn + i], xi);
      xs[i] = xi;
    }
    par ^= 1;
  }
  __syncthreads();
  double* out = Zt + ((size_t)z * V + v) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = xs[i];
}

// u = L^-T q for each eigenvector, one CTA per vector.  The diagonal blocks are applied through their
// precomputed inverses (Dinv from the Cholesky), the rest is a row-oriented update.  grid (V, nz).
__global__ void __launch_bounds__(BTT) eig_backsolve_kernel(const double* __restrict__ Lm,
                                                            const double* __restrict__ Dinv, double* __restrict__ Zt,
                                                            int n, int ldn, int V, int nblk) {
  extern __shared__ double xs[];
  __shared__ double us[NB];
  const int v = blockIdx.x, z = blockIdx.y;
  double* q = Zt + ((size_t)z * V + v) * n;
  const double* L = Lm + (size_t)z * n * ldn;
  const double* Di = Dinv + (size_t)z * nblk * NB * NB;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = q[i];
  __syncthreads();
  for (int b = nblk - 1; b >= 0; --b) {
    const int b0 = b * NB, bs = min(NB, n - b0);
    // u_b = Linv_bb^T q_b :  u[c] = sum_{r >= c} Linv[r][c] q[r]   (4 threads per output)
    if (threadIdx.x < 4 * NB) {            // warps 0..7
      const int c = threadIdx.x >> 2, part = threadIdx.x & 3;
      double acc = 0.0;
      if (c < bs)
        for (int r = c + part; r < bs; r += 4) acc = fma(Di[((size_t)b * NB + r) * NB + c], xs[b0 + r], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0 && c < bs) us[c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < b0; c += blockDim.x) {
      double acc = 0.0;
#pragma unroll 16
      for (int r = 0; r < bs; ++r) acc = fma(us[r], L[(size_t)(b0 + r) * ldn + c], acc);
      xs[c] -= acc;
    }
    if (threadIdx.x < bs) xs[b0 + threadIdx.x] = us[threadIdx.x];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] = xs[i];
}

// ------------------------------------------------------------------------------------------------
// Small-n path (eig_mode 2, automatic for n <= JACOBI_AUTO_N): one-sided (Hestenes) Jacobi on the symmetric C held
// in shared memory, one CTA per zone, one warp per column pair, round-robin parallel ordering.  Columns of G = C V
// are rotated until mutually orthogonal; V accumulates the rotations, so V holds the eigenvectors and
// lambda_i = v_i . g_i (sign included, no division -- rank-deficient C is fine).  Full spectrum, no tridiagonal
// stage, no clusters to repair; used where the matrix fits in one SM (2 n^2 doubles).
constexpr int JACOBI_MAX_N = 112;
constexpr int JACOBI_AUTO_N = 48;
constexpr int TWOSTAGE_AUTO_N = 1024;   // eig_mode 0: two-stage tridiagonalisation (band.cu) from this size on

__global__ void __launch_bounds__(256) eig_jacobi_kernel(const double* __restrict__ Cm, double* __restrict__ lam,
                                                         double* __restrict__ Zt, int* __restrict__ info, int n, int ldn,
                                                         int V) {
  extern __shared__ double jsm[];
  const int np = (n + 1) & ~1;
  double* G = jsm;                 // column-major np x np
  double* Vm = jsm + (size_t)np * np;
  __shared__ double lamv[JACOBI_MAX_N + 2];
  __shared__ int rotated;
  const int z = blockIdx.x;
  const double* C = Cm + (size_t)z * n * ldn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = threadIdx.x; e < np * np; e += blockDim.x) {
    const int c = e / np, r = e - c * np;
    G[e] = (r < n && c < n) ? C[(size_t)r * ldn + c] : 0.0;
    Vm[e] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int m1 = np - 1;
  int sweep = 0;
  for (; sweep < 40; ++sweep) {
    if (threadIdx.x == 0) rotated = 0;
    __syncthreads();
    for (int rd = 0; rd < m1; ++rd) {
      for (int k = warp; k < np / 2; k += nw) {
        int p, q;
        if (k == 0) { p = m1; q = rd; }
        else { p = (rd + k) % m1; q = (rd - k + m1) % m1; }
        double* gp = G + (size_t)p * np;
        double* gq = G + (size_t)q * np;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int i = lane; i < np; i += 32) {
          const double x = gp[i], y = gq[i];
          a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
        }
        a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
        if (fabs(g) > DBL_EPSILON * sqrt(a * b) && g != 0.0) {        // warp-uniform
          const double zeta = (b - a) / (2.0 * g);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double* vp = Vm + (size_t)p * np;
          double* vq = Vm + (size_t)q * np;
          for (int i = lane; i < np; i += 32) {
            const double x = gp[i], y = gq[i];
            gp[i] = cs * x - sn * y;
            gq[i] = sn * x + cs * y;
            const double u = vp[i], w = vq[i];
            vp[i] = cs * u - sn * w;
            vq[i] = sn * u + cs * w;
          }
          if (lane == 0) rotated = 1;
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  if (sweep >= 40 && threadIdx.x == 0) atomicOr(&info[z * 4 + 1], 2);
  for (int c = warp; c < n; c += nw) {
    double s = 0.0;
    for (int i = lane; i < np; i += 32) s = fma(Vm[(size_t)c * np + i], G[(size_t)c * np + i], s);
    s = warp_sum(s);
    if (lane == 0) lamv[c] = s;
  }
  __syncthreads();
  // rank sort, descending
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double lc = lamv[c];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (lamv[j] > lc) || (lamv[j] == lc && j < c);
    if (rank < V) {
      lam[(size_t)z * V + rank] = lc;
      double* out = Zt + ((size_t)z * V + rank) * n;
      for (int i = 0; i < n; ++i) out[i] = Vm[(size_t)c * np + i];
    }
  }
}

template <typename Kern>
int ensure_smem(Kern k, size_t bytes) {
  if (bytes > 48 * 1024) APV_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return OK;
}

// Zt[z][v][i] = Y[z][i][v]  (n x Vp row-major vectors-as-columns -> rows of Zt).  grid (ceil(V/32), ceil(n/32), nz), block (32, 8)
__global__ void vectors_to_rows_kernel(const double* __restrict__ Y, long long ystride, int Vp, double* __restrict__ Zt,
                                       int n, int V) {
  __shared__ double t[32][33];
  const int z = blockIdx.z;
  const double* y = Y + (size_t)z * ystride;
  double* o = Zt + (size_t)z * V * n;
  const int v0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, v = v0 + threadIdx.x;
    t[r][threadIdx.x] = (i < n && v < V) ? y[(size_t)i * Vp + v] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int v = v0 + r, i = i0 + threadIdx.x;
    if (v < V && i < n) o[(size_t)v * n + i] = t[threadIdx.x][r];
  }
}

}  // namespace

// =================================================================================================
int jdiag_alloc(JdiagWs& ws, int n, int V, int nz, int eig_mode) {
  ws.n = n; ws.V = V; ws.nz = nz; ws.nb = NB; ws.nbt = NBT; ws.eig_mode = eig_mode;
  ws.ldn = round_up(n, 8);
  ws.Vp = round_up(V, 32);
  const size_t mat = (size_t)nz * n * ws.ldn * sizeof(double);
  const int nblk = ceil_div(n, NB);
  size_t total = 0;
  auto al = [&](void** p, size_t bytes) -> int {
    APV_CUDA_TRY(cudaMalloc(p, bytes));
    APV_CUDA_TRY(cudaMemset(*p, 0, bytes));
    total += bytes;
    return OK;
  };
  APV_TRY(al((void**)&ws.Lm, mat));
  APV_TRY(al((void**)&ws.Cm, mat));
  APV_TRY(al((void**)&ws.Tm, mat));
  APV_TRY(al((void**)&ws.VH, mat));
  APV_TRY(al((void**)&ws.Dinv, (size_t)nz * nblk * NB * NB * sizeof(double)));
  APV_TRY(al((void**)&ws.SBinv, (size_t)nz * ceil_div(n, SB) * SB * SB * sizeof(double)));
  APV_TRY(al((void**)&ws.Z1, (size_t)nz * n * 2 * NBT * sizeof(double)));
  APV_TRY(al((void**)&ws.Z2, (size_t)nz * n * 2 * NBT * sizeof(double)));
  const size_t vec = (size_t)nz * n * sizeof(double);
  APV_TRY(al((void**)&ws.tau, vec));
  APV_TRY(al((void**)&ws.dd, vec));
  APV_TRY(al((void**)&ws.ee, vec));
  APV_TRY(al((void**)&ws.colbuf, vec));
  APV_TRY(al((void**)&ws.ybuf, vec));
  APV_TRY(al((void**)&ws.wbuf, vec));
  APV_TRY(al((void**)&ws.tdws, tridiag_scratch_doubles(n, nz) * sizeof(double)));
  APV_TRY(al((void**)&ws.Tf, (size_t)nz * ceil_div(n, WYB) * WYB * WYB * sizeof(double)));
  APV_TRY(al((void**)&ws.lam, (size_t)nz * V * sizeof(double)));
  APV_TRY(al((void**)&ws.shift, (size_t)nz * V * sizeof(double) + (size_t)nz * 4 * sizeof(double)));
  APV_TRY(al((void**)&ws.iv, (size_t)nz * 6 * n * ws.Vp * sizeof(double)));
  APV_TRY(al((void**)&ws.Zt, (size_t)nz * V * n * sizeof(double)));
  APV_TRY(al((void**)&ws.info, (size_t)nz * 4 * sizeof(int)));
  APV_TRY(al((void**)&ws.ts2, twostage_scratch_bytes(n, nz, twostage_nsplit_max())));
  for (auto& e : ws.ev) APV_CUDA_TRY(cudaEventCreate(&e));
  for (auto& e : ws.ev2) APV_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDefault));
  {
    int lo = 0, hi = 0;
    APV_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    APV_CUDA_TRY(cudaStreamCreateWithPriority(&ws.st2, cudaStreamNonBlocking, hi));
  }
  ws.npanel = ceil_div(n, NBT);
  ws.pev = new cudaEvent_t[2 * ws.npanel]();
  for (int i = 0; i < 2 * ws.npanel; ++i) APV_CUDA_TRY(cudaEventCreate(&ws.pev[i]));
  ws.bytes = total;
  return OK;
}

void jdiag_free(JdiagWs& ws) {
  void* ps[] = {ws.Lm, ws.Cm, ws.Tm, ws.VH, ws.Dinv, ws.Z1, ws.Z2, ws.tau, ws.dd, ws.ee, ws.colbuf,
                ws.ybuf, ws.wbuf, ws.tdws, ws.Tf, ws.lam, ws.shift, ws.iv, ws.Zt, ws.info, ws.ts2, ws.SBinv, ws.q1agg, ws.dcw};
  for (void* p : ps)
    if (p) cudaFree(p);
  for (auto& e : ws.ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : ws.ev2)
    if (e) cudaEventDestroy(e);
  if (ws.st2) cudaStreamDestroy(ws.st2);
  if (ws.pev) {
    for (int i = 0; i < 2 * ws.npanel; ++i)
      if (ws.pev[i]) cudaEventDestroy(ws.pev[i]);
    delete[] ws.pev;
  }
  ws = JdiagWs();
}

// Y <- L^-1 X for an n x m right-hand side (row-major, leading dimension ldx), batched over zones; X is destroyed.
// Blocked forward substitution on super-blocks of SB rows: Y_s = Linv_ss X_s with the explicit inverse of the
// diagonal super-block (sb_inv_kernel), then ONE rank-SB DMMA update of everything below, so the right-hand side
// is streamed n / SB times and there are two launches per super-block.
static int trsm_lower(JdiagWs& ws, double* X, double* Y, int m, int ldx, long long strideX, cudaStream_t st,
                      int* launches) {
  const int n = ws.n, ldn = ws.ldn, nsb = ceil_div(n, SB);
  const long long mstride = (long long)n * ldn;
  for (int s0 = 0; s0 < n; s0 += SB) {
    const int s1 = std::min(n, s0 + SB), rows = s1 - s0;
    GemmArgs g{};
    g.batch = ws.nz;
    g.A = ws.SBinv + (size_t)(s0 / SB) * SB * SB; g.lda = SB; g.strideA = (long long)nsb * SB * SB;
    g.B = X + (size_t)s0 * ldx; g.ldb = ldx; g.strideB = strideX;
    g.C = Y + (size_t)s0 * ldx; g.ldc = ldx; g.strideC = strideX;
    g.M = rows; g.N = m; g.K = rows; g.alpha = 1.0; g.beta = 0.0;
    APV_TRY(gemm_f64(g, st));
    ++*launches;
    if (s1 < n) {
      GemmArgs u{};
      u.batch = ws.nz;
      u.A = ws.Lm + (size_t)s1 * ldn + s0; u.lda = ldn; u.strideA = mstride;
      u.B = Y + (size_t)s0 * ldx; u.ldb = ldx; u.strideB = strideX;
      u.C = X + (size_t)s1 * ldx; u.ldc = ldx; u.strideC = strideX;
      u.M = n - s1; u.N = m; u.K = rows; u.alpha = -1.0; u.beta = 1.0;
      APV_TRY(gemm_f64(u, st));
      ++*launches;
    }
  }
  return OK;
}

// U = L^-T Q for MANY vectors (V > n / 8): blocked back substitution on super-blocks of SB rows with the explicit
// inverses of the diagonal super-blocks, bottom to top:  Y_s = Linv_ss^T X_s,  X[0:s0, :] -= L[s, 0:s0]^T Y_s.
// X (n x Vp, iv slot 4) is destroyed, Y (iv slot 0) receives U; all O(n^2 V) work is DMMA GEMM.
static int backsolve_gemm(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, Vp = ws.Vp, nsb = ceil_div(n, SB);
  double* X = ws.iv + 4 * (size_t)n * Vp;
  double* Y = ws.iv;
  const long long xs = 6LL * n * Vp, mstride = (long long)n * ldn;
  for (int s = nsb - 1; s >= 0; --s) {
    const int s0 = s * SB, s1 = std::min(n, s0 + SB), rows = s1 - s0;
    GemmArgs g{};
    g.batch = ws.nz;
    g.A = ws.SBinv + (size_t)s * SB * SB; g.lda = SB; g.strideA = (long long)nsb * SB * SB; g.transA = 1;
    g.B = X + (size_t)s0 * Vp; g.ldb = Vp; g.strideB = xs;
    g.C = Y + (size_t)s0 * Vp; g.ldc = Vp; g.strideC = xs;
    g.M = rows; g.N = Vp; g.K = rows; g.alpha = 1.0; g.beta = 0.0;
    APV_TRY(gemm_f64(g, st));
    ++*launches;
    if (s0 > 0) {
      GemmArgs u{};
      u.batch = ws.nz;
      u.A = ws.Lm + (size_t)s0 * ldn; u.lda = ldn; u.strideA = mstride; u.transA = 1;     // L[s0:s1, 0:s0]^T
      u.B = Y + (size_t)s0 * Vp; u.ldb = Vp; u.strideB = xs;
      u.C = X; u.ldc = Vp; u.strideC = xs;
      u.M = s0; u.N = Vp; u.K = rows; u.alpha = -1.0; u.beta = 1.0;
      APV_TRY(gemm_f64(u, st));
      ++*launches;
    }
  }
  return OK;
}

// U = L^-T Q for the top-V vectors kept as ROWS (Zt: V x n per zone):  X L = Z,  blocked from the last super-block to the
// first with the explicit inverses of the diagonal super-blocks,
//     X[:, s] = Z[:, s] Linv_ss,     Z[:, 0:s0] -= X[:, s] L[s, 0:s0],
// two small DMMA GEMMs per super-block (no transposes: both operands are read as stored).  X is built in the (free)
// inverse-iteration workspace and copied back.  The per-vector kernel (eig_backsolve_kernel: one CTA of 1024 threads per
// vector, 519 M warp instructions for 4.3 GFLOP at cfg-3) took 1.42 ms and 128 SMs.
static int backsolve_rows_gemm(JdiagWs& ws, cudaStream_t st, int* launches) {
  const int n = ws.n, ldn = ws.ldn, V = ws.V, nsb = ceil_div(n, SB);
  double* X = ws.iv;
  const long long zs = (long long)V * n, mstride = (long long)n * ldn;
  for (int s = nsb - 1; s >= 0; --s) {
    const int s0 = s * SB, s1 = std::min(n, s0 + SB), rows = s1 - s0;
    GemmArgs g{};
    g.batch = ws.nz;
    g.A = ws.Zt + s0; g.lda = n; g.strideA = zs;
    g.B = ws.SBinv + (size_t)s * SB * SB; g.ldb = SB; g.strideB = (long long)nsb * SB * SB;
    g.C = X + s0; g.ldc = n; g.strideC = zs;
    g.M = V; g.N = rows; g.K = rows; g.alpha = 1.0; g.beta = 0.0;
    APV_TRY(gemm_f64(g, st));
    ++*launches;
    if (s0 > 0) {
      GemmArgs u{};
      u.batch = ws.nz;
      u.A = X + s0; u.lda = n; u.strideA = zs;
      u.B = ws.Lm + (size_t)s0 * ldn; u.ldb = ldn; u.strideB = mstride;          // L[s0:s1, 0:s0]
      u.C = ws.Zt; u.ldc = n; u.strideC = zs;
      u.M = V; u.N = s0; u.K = rows; u.alpha = -1.0; u.beta = 1.0;
      APV_TRY(gemm_f64(u, st));
      ++*launches;
    }
  }
  APV_CUDA_TRY(cudaMemcpyAsync(ws.Zt, X, (size_t)ws.nz * V * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return OK;
}

int jdiag_run(JdiagWs& ws, const double* const bright[2], const double* const dark[2], int ld_in, double reg,
              cudaStream_t st, int* launches, const double* regv) {
  const int n = ws.n, ldn = ws.ldn, nz = ws.nz, V = ws.V;
  const long long mstride = (long long)n * ldn;
  int nl = 0;
  APV_CUDA_TRY(cudaEventRecord(ws.ev[0], st));
  APV_CUDA_TRY(cudaMemsetAsync(ws.info, 0, (size_t)nz * 4 * sizeof(int), st));
  Ptr2 pb, pd;
  for (int z = 0; z < 2; ++z) { pb.p[z] = bright[z < nz ? z : 0]; pd.p[z] = dark[z < nz ? z : 0]; }
  prep_kernel<<<dim3(ceil_div(n, 256), n, nz), 256, 0, st>>>(pb, pd, ld_in, ws.Cm, ws.Lm, n, ldn, reg, regv);
  ++nl;

  // ---- blocked Cholesky of Lm (lower)
  const int nblk = ceil_div(n, NB);
  const size_t chol_smem = (size_t)2 * NB * CDP * sizeof(double);
  APV_TRY(ensure_smem(chol_diag_kernel, chol_smem));
  // APV_CHOL_DEBUG=1: synchronous per-kernel-class times of the factorisation (diagonal blocks | panels | updates inside
  // a super-block | trailing updates), printed to stderr
  const bool cdbg = getenv("APV_CHOL_DEBUG") != nullptr;
  cudaEvent_t ce0 = nullptr, ce1 = nullptr;
  double cacc[4] = {0, 0, 0, 0};
  if (cdbg) { cudaEventCreate(&ce0); cudaEventCreate(&ce1); }
  auto cbegin = [&]() { if (cdbg) cudaEventRecord(ce0, st); };
  auto cend = [&](int k) {
    if (!cdbg) return;
    cudaEventRecord(ce1, st); cudaEventSynchronize(ce1);
    float ms = 0.f; cudaEventElapsedTime(&ms, ce0, ce1); cacc[k] += ms;
  };
  for (int s0 = 0; s0 < n; s0 += SB) {
    const int s1 = std::min(n, s0 + SB);
    for (int k0 = s0; k0 < s1; k0 += NB) {
      const int nbk = std::min(NB, n - k0);
      cbegin();
      chol_diag_kernel<<<nz, 128, chol_smem, st>>>(ws.Lm, ws.Dinv, n, ldn, k0, nbk, nblk, ws.info);
      cend(0);
      ++nl;
      const int below = n - k0 - nbk;
      if (below > 0) {
        GemmArgs p{};          // L[below, k] = A[below, k] * Linv_kk^T   (in place: one column tile)
        p.batch = nz;
        p.A = ws.Lm + (size_t)(k0 + nbk) * ldn + k0; p.lda = ldn; p.strideA = mstride;
        p.B = ws.Dinv + (size_t)(k0 / NB) * NB * NB; p.ldb = NB; p.strideB = (long long)nblk * NB * NB;
        p.C = ws.Lm + (size_t)(k0 + nbk) * ldn + k0; p.ldc = ldn; p.strideC = mstride;
        p.M = below; p.N = nbk; p.K = nbk; p.transB = 1; p.alpha = 1.0; p.beta = 0.0;
        cbegin();
        APV_TRY(gemm_f64(p, st));
        cend(1);
        ++nl;
        const int cols = s1 - k0 - nbk;      // remaining columns of the super-block
        if (cols > 0) {
          GemmArgs u{};        // A[below, sb cols] -= L[below, k] L[sb rows, k]^T
          u.batch = nz;
          u.A = p.C; u.lda = ldn; u.strideA = mstride;
          u.B = p.C; u.ldb = ldn; u.strideB = mstride;
          u.C = ws.Lm + (size_t)(k0 + nbk) * ldn + k0 + nbk; u.ldc = ldn; u.strideC = mstride;
          u.M = below; u.N = cols; u.K = nbk; u.transB = 1; u.alpha = -1.0; u.beta = 1.0;
          cbegin();
          APV_TRY(gemm_f64(u, st));
          cend(2);
          ++nl;
        }
      }
    }
    if (s1 < n) {
      GemmArgs u{};            // trailing: A22 -= L[rest, sb] L[rest, sb]^T  (lower tiles), K = super-block width
      u.batch = nz;
      u.A = ws.Lm + (size_t)s1 * ldn + s0; u.lda = ldn; u.strideA = mstride;
      u.B = u.A; u.ldb = ldn; u.strideB = mstride;
      u.C = ws.Lm + (size_t)s1 * ldn + s1; u.ldc = ldn; u.strideC = mstride;
      u.M = n - s1; u.N = n - s1; u.K = s1 - s0; u.transB = 1; u.tri = 1; u.alpha = -1.0; u.beta = 1.0;
      cbegin();
      APV_TRY(gemm_f64(u, st));
      cend(3);
      ++nl;
    }
  }
  if (cdbg) {
    fprintf(stderr, "cholesky ms: diagonal blocks %.2f | panels %.2f | super-block updates %.2f | trailing updates %.2f\n",
            cacc[0], cacc[1], cacc[2], cacc[3]);
    cudaEventDestroy(ce0); cudaEventDestroy(ce1);
  }

  APV_CUDA_TRY(cudaEventRecord(ws.ev[1], st));
  // ---- C = L^-1 A L^-T, symmetrised
  {
    const size_t sism = (size_t)3 * NB * (NB + 1) * sizeof(double);
    APV_TRY(ensure_smem(sb_inv_kernel, sism));
    sb_inv_kernel<<<dim3(ceil_div(n, SB), nz), 256, sism, st>>>(ws.Lm, ws.Dinv, ws.SBinv, n, ldn, nblk);
    ++nl;
  }
  APV_TRY(trsm_lower(ws, ws.Cm, ws.Tm, n, ldn, mstride, st, &nl));          // Tm = X = L^-1 A (all of it: n^3)
  dim3 tb(32, 8), tg(ceil_div(n, 32), ceil_div(n, 32), nz);
  if (getenv("APV_REDUCE_FULL")) {          // round-1 form: a second full solve around a transpose (2 n^3 in all)
    transpose_kernel<<<tg, tb, 0, st>>>(ws.Tm, ws.Cm, n, ldn);                // Cm = A L^-T
    ++nl;
    APV_TRY(trsm_lower(ws, ws.Cm, ws.Tm, n, ldn, mstride, st, &nl));          // Tm = L^-1 A L^-T
    symmetrize_kernel<<<tg, tb, 0, st>>>(ws.Tm, ws.Cm, n, ldn);
    ++nl;
  } else {
    // C = X L^-T is symmetric: only its lower block triangle is formed, column super-block by column super-block
    //   C[s0:, s] = X[s0:, s] Linv_ss^T,   X[s1:, s1:] -= C[s1:, s] L[s1:, s]^T  (lower tiles only),
    // n^3 / 3 instead of n^3 flops and no transpose; the upper triangle is mirrored.
    const int nsb = ceil_div(n, SB);
    for (int s0 = 0; s0 < n; s0 += SB) {
      const int s1 = std::min(n, s0 + SB), rows = s1 - s0;
      GemmArgs g{};
      g.batch = nz;
      g.A = ws.Tm + (size_t)s0 * ldn + s0; g.lda = ldn; g.strideA = mstride;
      g.B = ws.SBinv + (size_t)(s0 / SB) * SB * SB; g.ldb = SB; g.strideB = (long long)nsb * SB * SB; g.transB = 1;
      g.C = ws.Cm + (size_t)s0 * ldn + s0; g.ldc = ldn; g.strideC = mstride;
      g.M = n - s0; g.N = rows; g.K = rows; g.alpha = 1.0; g.beta = 0.0;
      APV_TRY(gemm_f64(g, st));
      ++nl;
      if (s1 < n) {
        GemmArgs u{};
        u.batch = nz;
        u.A = ws.Cm + (size_t)s1 * ldn + s0; u.lda = ldn; u.strideA = mstride;
        u.B = ws.Lm + (size_t)s1 * ldn + s0; u.ldb = ldn; u.strideB = mstride; u.transB = 1;
        u.C = ws.Tm + (size_t)s1 * ldn + s1; u.ldc = ldn; u.strideC = mstride;
        u.M = n - s1; u.N = n - s1; u.K = rows; u.alpha = -1.0; u.beta = 1.0; u.tri = 1;
        APV_TRY(gemm_f64(u, st));
        ++nl;
      }
    }
    mirror_lower_kernel<<<tg, tb, 0, st>>>(ws.Cm, n, ldn);
    ++nl;
  }

  APV_CUDA_TRY(cudaEventRecord(ws.ev[2], st));
  // ---- blocked Householder tridiagonalisation of Cm
  const bool use_jacobi = (ws.eig_mode == 2 && n <= JACOBI_MAX_N) || (ws.eig_mode == 0 && n <= JACOBI_AUTO_N);
  if (use_jacobi) {
    ws.last_two_stage = ws.last_panels = false;
    const int np = (n + 1) & ~1;
    const size_t jsm = (size_t)2 * np * np * sizeof(double);
    APV_TRY(ensure_smem(eig_jacobi_kernel, jsm));
    eig_jacobi_kernel<<<nz, 256, jsm, st>>>(ws.Cm, ws.lam, ws.Zt, ws.info, n, ldn, V);
    ++nl;
    for (int e = 3; e <= 5; ++e) APV_CUDA_TRY(cudaEventRecord(ws.ev[e], st));
    APV_TRY(ensure_smem(eig_backsolve_kernel, (size_t)n * sizeof(double)));
    eig_backsolve_kernel<<<dim3(V, nz), BTT, (size_t)n * sizeof(double), st>>>(ws.Lm, ws.Dinv, ws.Zt, n, ldn, V, nblk);
    ++nl;
    APV_CUDA_TRY(cudaEventRecord(ws.ev[6], st));
    APV_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += nl;
    return OK;
  }
  const bool two_stage = ws.eig_mode == 3 || (ws.eig_mode == 0 && n >= TWOSTAGE_AUTO_N);
  ws.last_two_stage = two_stage;
  ws.last_panels = !two_stage;
  if (two_stage) APV_TRY(twostage_run(ws, st, &nl));
  else APV_TRY(tridiag_run(ws, st, &nl));
  APV_CUDA_TRY(cudaEventRecord(ws.ev[3], st));
  // ---- eigenpairs of T
  // many vectors (full-spectrum requests): one LANE per vector fills the chip; few vectors: one CTA per vector with the
  // work arrays in shared memory (one CTA per SM, V / 148 waves of ~2.3 ms at n = 4096)
  const bool many = (long long)V * nz >= 2048 && !getenv("APV_EIG_FEW");
  // the whole spectrum of a two-stage problem: divide and conquer (dc.cu), its O(n^3) part on the tensor cores
  const bool use_dc = many && two_stage && V == n && ws.Vp >= n && dc_supported(n) && !getenv("APV_EIG_NO_DC");
  double* tnorm = ws.shift + (size_t)nz * V;
  if (use_dc) {
    APV_TRY(dc_run(ws, st, &nl));
  } else {
  if (V * nz <= 512) {
    APV_TRY(ensure_smem(eig_bisect_kernel<BIS_T>, (size_t)2 * n * sizeof(double)));
    eig_bisect_kernel<BIS_T><<<dim3(V, nz), BIS_T, (size_t)2 * n * sizeof(double), st>>>(ws.dd, ws.ee, ws.lam, tnorm, n, V);
  } else {
    APV_TRY(ensure_smem(eig_bisect_kernel<32>, (size_t)2 * n * sizeof(double)));
    eig_bisect_kernel<32><<<dim3(ceil_div(V, 8), nz), BIS_T, (size_t)2 * n * sizeof(double), st>>>(ws.dd, ws.ee, ws.lam, tnorm, n, V);
  }
  eig_shift_kernel<<<nz, 32, 0, st>>>(ws.lam, ws.shift, V);
  const size_t ivsm = (size_t)6 * n * sizeof(double) + (size_t)round_up(n, 16);
  if (ivsm <= 220 * 1024 && !many) {
    APV_TRY(ensure_smem(eig_invit_smem_kernel, ivsm));
    eig_invit_smem_kernel<<<dim3(V, nz), 128, ivsm, st>>>(ws.dd, ws.ee, ws.shift, tnorm, ws.iv, ws.info, n, V, ws.Vp);
  } else {
    eig_invit_kernel<<<dim3(ws.Vp / 32, nz), 32, 0, st>>>(ws.dd, ws.ee, ws.shift, tnorm, ws.iv, ws.info, n, V, ws.Vp);
  }
  eig_cluster_mgs_kernel<<<nz, 256, 0, st>>>(ws.lam, tnorm, ws.iv, n, V, ws.Vp, 1e-6);
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev[4], st));
  const bool gemm_bt = two_stage && many && n >= 512;
  if (gemm_bt) {
    // full-spectrum back-transformation on the tensor cores: Q2 wavefront, Q1 and U = L^-T Q as GEMMs on the n x Vp
    // matrix of vectors, then one transpose into the rows of Zt
    APV_TRY(twostage_apply_q2(ws, st, &nl));
    APV_TRY(twostage_apply_q1_gemm(ws, st, &nl));
    APV_CUDA_TRY(cudaEventRecord(ws.ev[5], st));
    APV_TRY(backsolve_gemm(ws, st, &nl));
    vectors_to_rows_kernel<<<dim3(ceil_div(V, 32), ceil_div(n, 32), nz), dim3(32, 8), 0, st>>>(ws.iv, 6LL * n * ws.Vp, ws.Vp, ws.Zt, n, V);
    nl += 6;
    APV_CUDA_TRY(cudaEventRecord(ws.ev[6], st));
    APV_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += nl;
    return OK;
  }
  if (two_stage) {
    APV_TRY(twostage_apply_q2(ws, st, &nl));
    APV_TRY(twostage_apply_q1(ws, st, &nl));
  } else {
    APV_TRY(ensure_smem(eig_backtransform_kernel, (size_t)n * sizeof(double)));
    wy_tfactor_kernel<<<dim3(ceil_div(n, WYB), nz), 256, 0, st>>>(ws.VH, ws.tau, ws.Tf, n, ldn);
    eig_backtransform_kernel<<<dim3(V, nz), BTT, (size_t)n * sizeof(double), st>>>(ws.iv, ws.VH, ws.Tf, ws.Zt, n, ldn, V, ws.Vp);
    ++nl;
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev[5], st));
  // the workspace of the inverse iteration (6 n Vp doubles per zone) holds the V x n result of the GEMM form
  static const bool bs_kernel = getenv("APV_BACKSOLVE_KERNEL") != nullptr;
  if (!bs_kernel && V >= 16 && n >= 2 * SB && (size_t)V * n <= (size_t)6 * n * ws.Vp) {
    APV_TRY(backsolve_rows_gemm(ws, st, &nl));
    nl += 5;
  } else {
    APV_TRY(ensure_smem(eig_backsolve_kernel, (size_t)n * sizeof(double)));
    eig_backsolve_kernel<<<dim3(V, nz), BTT, (size_t)n * sizeof(double), st>>>(ws.Lm, ws.Dinv, ws.Zt, n, ldn, V, nblk);
    nl += 6;
  }
  APV_CUDA_TRY(cudaEventRecord(ws.ev[6], st));
  APV_CUDA_TRY(cudaGetLastError());
  if (launches) *launches += nl;
  return OK;
}

}  // namespace apv
