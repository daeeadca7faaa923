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
