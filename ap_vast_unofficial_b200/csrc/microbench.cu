// FP64 tensor-core (DMMA.8x8x4) issue-rate microbenchmark: the roofline denominator for the
// statistics SYRK and the BLAS-3 parts of the joint diagonalisation (SURVEY.md section 2.2 note:
// B200's FP64 tensor peak is not in MEASURED_PEAKS.json and has to be measured on the box).
#include "engine.cuh"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a0, double b0) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) apv::dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int apv_bench_dmma_peak(int iters, double* tflops) {
  using namespace apv;
  if (iters < 1 || !tflops) return EINVAL_;
  int dev = 0, sms = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 4;
  double* out = nullptr;
  APV_CUDA_TRY(cudaMalloc((void**)&out, (size_t)grid * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dmma_peak_kernel<<<grid, 256>>>(out, iters / 10 + 1, 1.0, 0.5);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, 0);
    dmma_peak_kernel<<<grid, 256>>>(out, iters, 1.0, 0.5);
    cudaEventRecord(e1, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "dmma peak kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      cudaFree(out);
      return ECUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // per warp and iteration: 16 DMMA of 8x8x4 -> 16 * 2*8*8*4 flops
    const double flops = (double)grid * 8.0 * (double)iters * 16.0 * 512.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return OK;
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA (CUDA-core DFMA) throughput and dependent-issue latency: decides how much scalar FP64 the
// latency-bound kernels (panel QR, bulge chasing, inverse iteration) can afford next to the DMMA pipe.
namespace {

__global__ void __launch_bounds__(256) dfma_rate_kernel(double* out, int iters, double a0, double b0) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i * 1e-3;
  const double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_latency_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double x = a;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x = fma(x, a, b);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

}  // namespace

// out3: [0] DFMA TFLOP/s of the whole chip (16 independent chains per thread, 4 x 256 threads per SM),
//       [1] cycles per dependent DFMA (one warp), [2] warp-level DFMA issue interval in cycles per SM sub-partition
extern "C" int apv_bench_dfma(int iters, double* out3) {
  using namespace apv;
  if (iters < 1 || !out3) return EINVAL_;
  int dev = 0, sms = 0, khz = 0;
  APV_CUDA_TRY(cudaGetDevice(&dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  APV_CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const int grid = sms * 4;
  double* out = nullptr;
  long long* cyc = nullptr;
  APV_CUDA_TRY(cudaMalloc((void**)&out, (size_t)grid * 256 * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&cyc, sizeof(long long)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_rate_kernel<<<grid, 256>>>(out, iters / 10 + 1, 1.0, 0.5);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, 0);
    dfma_rate_kernel<<<grid, 256>>>(out, iters, 0.999, 0.5);
    cudaEventRecord(e1, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "dfma kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      return ECUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = (double)grid * 256.0 * (double)iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  dfma_latency_kernel<<<1, 32>>>(out, cyc, iters, 0.999, 0.5);
  long long hc = 0;
  APV_CUDA_TRY(cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
  out3[0] = best;
  out3[1] = (double)hc / ((double)iters * 16.0);
  // FMA/clk/SM = TFLOP/s / 2 / (SMs * clock); a warp instruction is 32 FMAs on one of 4 sub-partitions
  const double fma_per_clk_sm = best * 1e12 / 2.0 / ((double)sms * (double)khz * 1e3);
  out3[2] = 32.0 / (fma_per_clk_sm / 4.0);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaFree(cyc);
  return OK;
}
