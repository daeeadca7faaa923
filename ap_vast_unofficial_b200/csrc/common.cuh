// Shared device/host helpers for the AP-VAST B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

namespace apv {

// ----------------------------------------------------------------------------------------------
// status codes (mirrored in include/apvast_b200.h)
enum { OK = 0, EINVAL_ = 1, ENOTPD = 2, ECUDA = 3, ENOMEM_ = 4, ENOCONV = 5, ENCCL = 6 };

#define APV_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      snprintf(apv::g_err, sizeof(apv::g_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,   \
               cudaGetErrorString(_e));                                                        \
      return apv::ECUDA;                                                                       \
    }                                                                                          \
  } while (0)

#define APV_TRY(expr)                  \
  do {                                 \
    int _s = (expr);                   \
    if (_s != apv::OK) return _s;      \
  } while (0)

extern thread_local char g_err[512];

// NVTX range around the ENQUEUEING of a stage (header-only NVTX 3: a no-op unless a profiler is attached); the per-stage
// device times come from CUDA events (apv_stage_times), the ranges label the launches in ncu / nsys timelines.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// cudaFuncSetAttribute belongs to the device context, not to the calling thread: bookkeeping per device ordinal.
struct PerDevice {
  size_t v[64] = {};
  size_t& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return v[d & 63];
  }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ----------------------------------------------------------------------------------------------
// FP64 tensor-core atom.  On sm_100a the only FP64 MMA is the legacy warp-synchronous
// mma.sync m8n8k4 (SASS DMMA.8x8x4); tcgen05 has no f64 kind.
//   A (8x4, row): lane holds A[lane>>2][lane&3]
//   B (4x8, col): lane holds B[lane&3][lane>>2]
//   C (8x8)     : lane holds C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; every thread gets the result.  `red` is shared scratch of >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ----------------------------------------------------------------------------------------------
// GEMM (gemm.cu).  Row-major everywhere.  C[M x N] = alpha * op(A) * op(B) + beta * C.
//   transA == 0: A is M x K (lda);  transA == 1: A is stored K x M (lda)  (op(A) = A^T)
//   transB == 0: B is K x N (ldb);  transB == 1: B is stored N x K (ldb)  (op(B) = B^T)
//   tri: 0 = all tiles; 1 = only tiles that touch the lower triangle (row >= col) of C;
//        (elements above the diagonal inside diagonal tiles are still written).
//   batch: operands advance by strideA/B/C elements per batch entry;
//   split (> 1): every batch entry holds `split` sub-products (e.g. K slices writing partial results) whose operands
//        advance by splitA/B/C elements; blockIdx.z = batch index * split + sub-product.  With split_ktot the
//        sub-products are slices of one K range of that length (the last slice may be short or empty).
struct GemmArgs {
  const double* A; const double* B; double* C;
  int M, N, K;
  int lda, ldb, ldc;
  long long strideA, strideB, strideC;
  double alpha, beta;
  int transA, transB, tri, batch;
  int split;
  long long splitA, splitB, splitC;
  int split_ktot;   // > 0: sub-product s covers K indices [s K, min((s + 1) K, split_ktot))
  int mirror;       // with tri (square C, symmetric result): tiles strictly below the diagonal are also stored transposed
  int bn;           // 0: tile width by N (32 / 64 / 128); 64: force the 128 x 64 tile (two CTAs per SM: short-K updates)
};
int gemm_f64(const GemmArgs& g, cudaStream_t st);

}  // namespace apv
