// FP64 tensor-core (DMMA m8n8k4) GEMM building block used by the joint-diagonalisation kernels
// (blocked Cholesky trailing update, blocked triangular solves, SYR2K of the tridiagonalisation,
// back-transformations).  Replaces the BLAS-3 calls underneath the reference's jdiag
// (Python/apvast.py:20-36: LAPACK dpotrf/dtrtrs + BLAS dgemm/dsyrk).
//
// CTA tile 128 x BN x 16 (BN = 128, 64 or 32 by the width of the product), 8 warps, warp tile (BM / warps_m) x 32:
// 64x32 = 8x4 DMMA atoms (64 FP64 accumulators/thread) at BN = 128, 32x32 at BN = 64, 16x32 at BN = 32 (the
// skinny products of the band reduction, which are bandwidth-bound and run two CTAs per SM).  Register-staged
// double buffering of the global->shared copies, padded shared tiles so every fragment load is bank-conflict
// free (row pitch == 4 mod 16 doubles).  blockIdx.z = batch * split + s: `split` independent K-slices or
// sub-problems inside one batch entry advance the operands by splitA/B/C elements.
#include <stdlib.h>

#include "common.cuh"

namespace apv {

thread_local char g_err[512] = {0};

namespace {

constexpr int BM = 128, BK = 16;
constexpr int LDK = BK + 4;    // pitch of [rows][BK] tiles      (20  == 4 mod 16)
constexpr int LDM = BM + 4;    // pitch of [BK][BM] tiles        (132 == 4 mod 16)
constexpr int TILE_A = (BM * LDK > BK * LDM) ? BM * LDK : BK * LDM;   // doubles per A tile
__host__ __device__ constexpr int tile_b(int BN) { return (BN * LDK > BK * (BN + 4)) ? BN * LDK : BK * (BN + 4); }
constexpr int smem_bytes(int BN) { return 2 * (TILE_A + tile_b(BN)) * (int)sizeof(double); }   // 2 stages

// Fetch 8 consecutive doubles of a row-major matrix with bounds (zero fill).
__device__ __forceinline__ void fetch8(const double* __restrict__ P, int ld, int r, int c, int rmax, int cmax,
                                       bool vec, double (&v)[8]) {
  if (r < rmax && c + 7 < cmax && vec) {
    const double2* p = reinterpret_cast<const double2*>(P + (size_t)r * ld + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double2 t = __ldg(p + i);
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (r < rmax && c + i < cmax) ? __ldg(P + (size_t)r * ld + c + i) : 0.0;
  }
}

template <int TA, int TB, int BN>
__global__ void __launch_bounds__(256, BN == 128 ? 1 : 2) gemm_kernel(GemmArgs g) {
  constexpr int WN = BN / 32, WM = 8 / WN, WR = BM / WM, RT = WR / 8;   // warp grid, warp rows, row atoms
  constexpr int LDN = BN + 4;                                           // pitch of [BK][BN] tiles (== 4 mod 16)
  constexpr int TILE_B = tile_b(BN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (g.tri && n0 > m0 + BM - 1) return;
  extern __shared__ __align__(16) double smem[];
  double* As[2] = {smem, smem + TILE_A};
  double* Bs[2] = {smem + 2 * TILE_A, smem + 2 * TILE_A + TILE_B};

  const int nsp = g.split > 1 ? g.split : 1;
  const int zb = blockIdx.z / nsp, zs = blockIdx.z % nsp;
  const double* __restrict__ A = g.A + (size_t)zb * g.strideA + (size_t)zs * g.splitA;
  const double* __restrict__ B = g.B + (size_t)zb * g.strideB + (size_t)zs * g.splitB;
  double* __restrict__ C = g.C + (size_t)zb * g.strideC + (size_t)zs * g.splitC;
  const int K = g.split_ktot > 0 ? max(0, min(g.K, g.split_ktot - zs * g.K)) : g.K;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp / WN) * WR, wn = (warp % WN) * 32;
  const int gq = lane >> 2, tq = lane & 3;
  const bool vecA = ((g.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = ((g.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  // global->shared mapping
  //  [rows][BK] tiles: thread -> row tid>>1, cols (tid&1)*8 .. +7
  //  [BK][rows] tiles: thread -> k-row tid / (rows/8), cols (tid % (rows/8))*8 .. +7
  // (the B tile has BN rows / columns: only the first 2 BN threads carry a piece of it)
  const int rA = TA ? (tid >> 4) : (tid >> 1), cA = TA ? (tid & 15) * 8 : (tid & 1) * 8;
  const int rB = TB ? (tid >> 1) : (tid / (BN / 8)), cB = TB ? (tid & 1) * 8 : (tid % (BN / 8)) * 8;
  const bool hasB = tid < 2 * BN;

  double ra[8], rb[8];
  auto gload = [&](int k0) {
    if (TA) fetch8(A, g.lda, k0 + rA, m0 + cA, K, g.M, vecA, ra);
    else    fetch8(A, g.lda, m0 + rA, k0 + cA, g.M, K, vecA, ra);
    if (hasB) {
      if (TB) fetch8(B, g.ldb, n0 + rB, k0 + cB, g.N, K, vecB, rb);
      else    fetch8(B, g.ldb, k0 + rB, n0 + cB, K, g.N, vecB, rb);
    }
  };
  auto sstore = [&](int s) {
    double* a = As[s] + rA * (TA ? LDM : LDK) + cA;
    double* b = Bs[s] + rB * (TB ? LDK : LDN) + cB;
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(a)[i] = make_double2(ra[2 * i], ra[2 * i + 1]);
    if (hasB) {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<double2*>(b)[i] = make_double2(rb[2 * i], rb[2 * i + 1]);
    }
  };

  double acc[RT][4][2];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (K + BK - 1) / BK;
  if (nk > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
    const double* a_s = As[s];
    const double* b_s = Bs[s];
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[RT], b[4];
#pragma unroll
      for (int rt = 0; rt < RT; ++rt)
        a[rt] = TA ? a_s[(kk * 4 + tq) * LDM + wm + rt * 8 + gq] : a_s[(wm + rt * 8 + gq) * LDK + kk * 4 + tq];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
        b[ct] = TB ? b_s[(wn + ct * 8 + gq) * LDK + kk * 4 + tq] : b_s[(kk * 4 + tq) * LDN + wn + ct * 8 + gq];
#pragma unroll
      for (int rt = 0; rt < RT; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
    }
    if (kt + 1 < nk) sstore(s ^ 1);
    __syncthreads();
  }

  const bool vecC = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  const bool mir = g.tri && g.mirror && m0 >= n0 + BN;
#pragma unroll
  for (int rt = 0; rt < RT; ++rt) {
    const int r = m0 + wm + rt * 8 + gq;
    if (r >= g.M) continue;
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      const int c = n0 + wn + ct * 8 + 2 * tq;
      double* p = C + (size_t)r * g.ldc + c;
      double v0 = g.alpha * acc[rt][ct][0], v1 = g.alpha * acc[rt][ct][1];
      if (c + 1 < g.N && vecC) {
        if (g.beta != 0.0) {
          double2 o = *reinterpret_cast<double2*>(p);
          v0 += g.beta * o.x;
          v1 += g.beta * o.y;
        }
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
      } else {
        if (c < g.N) p[0] = v0 = v0 + (g.beta != 0.0 ? g.beta * p[0] : 0.0);
        if (c + 1 < g.N) p[1] = v1 = v1 + (g.beta != 0.0 ? g.beta * p[1] : 0.0);
      }
      if (mir) {          // 8 lanes (gq) write 8 consecutive doubles of row c: full 32-byte sectors
        if (c < g.N) C[(size_t)c * g.ldc + r] = v0;
        if (c + 1 < g.N) C[(size_t)(c + 1) * g.ldc + r] = v1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cp.async variant of the 128 x 64 and 128 x 32 tiles for op(A) = A (row-major M x K) with 16-byte aligned rows:
// the operand tiles are staged by 16-byte cp.async copies (zero fill at the edges) through a ring of shared-memory
// stages, so a CTA keeps STAGES - 1 k-steps in flight instead of one register-staged step; two CTAs per SM.
// The band reduction's Y = C22 V (N = 32) is a STREAM of C22 (8 flops per byte: 1.8 TB/s register-staged, 6.8 ->
// 4.9 ms with this kernel); the rank-64 .. rank-256 updates of the Cholesky, the triangular solves and the band
// reduction run 16 k-steps or fewer per tile, where the load latency of every step was exposed.
constexpr int AS_LDA = BK + 4;                                   // pitch of [rows][BK] tiles (== 4 mod 16)
template <int TB, int BN> struct AsyncCfg {
  static constexpr int LDB = TB ? BK + 4 : BN + 4;
  static constexpr int B_DOUBLES = TB ? BN * (BK + 4) : BK * (BN + 4);
  static constexpr int STAGE = BM * AS_LDA + B_DOUBLES;
  static constexpr int STAGES = BN == 32 ? 4 : 3;
  static constexpr int SMEM = STAGES * STAGE * (int)sizeof(double);
};

__device__ __forceinline__ void cp_async16(double* dst, const double* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int TB, int BN>
__global__ void __launch_bounds__(256, 2) gemm_async_kernel(GemmArgs g) {
  using Cfg = AsyncCfg<TB, BN>;
  constexpr int WN = BN / 32, WM = 8 / WN, WR = BM / WM, RT = WR / 8;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) double smem[];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (g.tri && n0 > m0 + BM - 1) return;
  const int nsp = g.split > 1 ? g.split : 1;
  const int zb = blockIdx.z / nsp, zs = blockIdx.z % nsp;
  const double* __restrict__ A = g.A + (size_t)zb * g.strideA + (size_t)zs * g.splitA;
  const double* __restrict__ B = g.B + (size_t)zb * g.strideB + (size_t)zs * g.splitB;
  double* __restrict__ C = g.C + (size_t)zb * g.strideC + (size_t)zs * g.splitC;
  const int K = g.split_ktot > 0 ? max(0, min(g.K, g.split_ktot - zs * g.K)) : g.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp / WN) * WR, wn = (warp % WN) * 32;
  const int gq = lane >> 2, tq = lane & 3;
  const int nk = (K + BK - 1) / BK;

  auto issue = [&](int kt) {
    double* as = smem + (size_t)(kt % STAGES) * Cfg::STAGE;
    double* bs = as + BM * AS_LDA;
    const int k0 = kt * BK;
#pragma unroll
    for (int u = 0; u < 4; ++u) {           // A tile: 128 rows x 8 chunks of 2 doubles
      const int c = tid + 256 * u, row = c >> 3, col = k0 + 2 * (c & 7);
      const bool ok = m0 + row < g.M && col < K;
      cp_async16(as + row * AS_LDA + 2 * (c & 7), ok ? A + (size_t)(m0 + row) * g.lda + col : A,
                 ok ? min(16, (K - col) * 8) : 0);
    }
    if (TB) {                                // B stored N x K: BN rows x 8 chunks
#pragma unroll
      for (int u = 0; u < BN / 32; ++u) {
        const int c = tid + 256 * u, row = c >> 3, col = k0 + 2 * (c & 7);
        const bool ok = n0 + row < g.N && col < K;
        cp_async16(bs + row * Cfg::LDB + 2 * (c & 7), ok ? B + (size_t)(n0 + row) * g.ldb + col : B,
                   ok ? min(16, (K - col) * 8) : 0);
      }
    } else {                                 // B stored K x N: 16 rows x BN / 2 chunks
#pragma unroll
      for (int u = 0; u < BN / 32; ++u) {
        const int c = tid + 256 * u, krow = c / (BN / 2), col = 2 * (c % (BN / 2));
        const bool ok = k0 + krow < K && n0 + col < g.N;
        cp_async16(bs + krow * Cfg::LDB + col, ok ? B + (size_t)(k0 + krow) * g.ldb + n0 + col : B,
                   ok ? min(16, (g.N - n0 - col) * 8) : 0);
      }
    }
  };

  double acc[RT][4][2];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) issue(s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (g.beta != 0.0) {
    // read-modify-write epilogue: the C tile is pulled into L2 while the K loop runs.  With 128 registers per thread the
    // epilogue cannot keep all its loads in flight, so on a short K loop it paid several DRAM round trips one after the
    // other (rank-64 band update at n = 4032: 0.155 ms with beta = 1 against 0.093 ms with beta = 0).
    constexpr int LPR = BN / 16;                       // 128-byte lines per tile row
    for (int i = tid; i < BM * LPR; i += 256) {
      const int row = m0 + i / LPR, col = n0 + (i % LPR) * 16;
      if (row < g.M && col < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (size_t)row * g.ldc + col));
    }
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    __syncthreads();
    if (kt + STAGES - 1 < nk) issue(kt + STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* a_s = smem + (size_t)(kt % STAGES) * Cfg::STAGE;
    const double* b_s = a_s + BM * AS_LDA;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[RT], b[4];
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) a[rt] = a_s[(wm + rt * 8 + gq) * AS_LDA + kk * 4 + tq];
#pragma unroll
      for (int ct = 0; ct < 4; ++ct)
        b[ct] = TB ? b_s[(wn + ct * 8 + gq) * Cfg::LDB + kk * 4 + tq] : b_s[(kk * 4 + tq) * Cfg::LDB + wn + ct * 8 + gq];
#pragma unroll
      for (int rt = 0; rt < RT; ++rt)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt], b[ct]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  const bool vecC = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  const bool mir = g.tri && g.mirror && m0 >= n0 + BN;
#pragma unroll
  for (int rt = 0; rt < RT; ++rt) {
    const int r = m0 + wm + rt * 8 + gq;
    if (r >= g.M) continue;
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      const int c = n0 + wn + ct * 8 + 2 * tq;
      double* p = C + (size_t)r * g.ldc + c;
      double v0 = g.alpha * acc[rt][ct][0], v1 = g.alpha * acc[rt][ct][1];
      if (c + 1 < g.N && vecC) {
        if (g.beta != 0.0) {
          double2 o = *reinterpret_cast<double2*>(p);
          v0 += g.beta * o.x;
          v1 += g.beta * o.y;
        }
        *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
      } else {
        if (c < g.N) p[0] = v0 = v0 + (g.beta != 0.0 ? g.beta * p[0] : 0.0);
        if (c + 1 < g.N) p[1] = v1 = v1 + (g.beta != 0.0 ? g.beta * p[1] : 0.0);
      }
      if (mir) {
        if (c < g.N) C[(size_t)c * g.ldc + r] = v0;
        if (c + 1 < g.N) C[(size_t)(c + 1) * g.ldc + r] = v1;
      }
    }
  }
}

template <int TB, int BN>
int launch_async(const GemmArgs& g, cudaStream_t st) {
  using Cfg = AsyncCfg<TB, BN>;
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (!configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(gemm_async_kernel<TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), (g.batch > 0 ? g.batch : 1) * (g.split > 1 ? g.split : 1));
  gemm_async_kernel<TB, BN><<<grid, 256, Cfg::SMEM, st>>>(g);
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

template <int TA, int TB, int BN>
int launch_bn(const GemmArgs& g, cudaStream_t st) {
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (!configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes(BN)));
    configured = true;
  }
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), (g.batch > 0 ? g.batch : 1) * (g.split > 1 ? g.split : 1));
  gemm_kernel<TA, TB, BN><<<grid, 256, smem_bytes(BN), st>>>(g);
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

template <int TA, int TB>
int launch(const GemmArgs& g, cudaStream_t st) {
  if (g.N <= 32) return launch_bn<TA, TB, 32>(g, st);
  // 128 x 64 tiles run two CTAs per SM: better for every product of this engine with K <= ~1000 (the delayed
  // rank-64 .. rank-256 updates are latency-bound per tile); the 128 x 128 tile is kept for long-K products
  if (g.N <= 64 || g.bn == 64 || (g.bn == 0 && g.K <= 1024)) return launch_bn<TA, TB, 64>(g, st);
  return launch_bn<TA, TB, 128>(g, st);
}

}  // namespace

int gemm_f64(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return OK;
  // cp.async tiles: op(A) = A with 16-byte aligned rows of both operands, the 128 x 64 / 128 x 32 tile range
  static const bool no_async = getenv("APV_GEMM_NO_ASYNC") != nullptr;
  // (every eligible product: the 128 x 64 cp.async tile also beats the register-staged 128 x 128 tile on long K --
  // 4096^3 and K = 2048, B transposed: 32.9 against 25.5 TFLOP/s, 89 % against 69 % of the DMMA peak)
  const bool tile64 = true;
  if (!no_async && !g.transA && tile64 && g.K >= 2 * BK && (g.lda & 1) == 0 && (g.ldb & 1) == 0 &&
      (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 &&
      (g.strideA & 1) == 0 && (g.strideB & 1) == 0 && (g.splitA & 1) == 0 && (g.splitB & 1) == 0) {
    if (g.N <= 32) return g.transB ? launch_async<1, 32>(g, st) : launch_async<0, 32>(g, st);
    return g.transB ? launch_async<1, 64>(g, st) : launch_async<0, 64>(g, st);
  }
  if (g.transA) return g.transB ? launch<1, 1>(g, st) : launch<1, 0>(g, st);
  return g.transB ? launch<0, 1>(g, st) : launch<0, 0>(g, st);
}

}  // namespace apv
