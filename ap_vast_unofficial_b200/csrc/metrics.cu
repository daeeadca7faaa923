// Evaluation metrics of the callers' side of the path: predicted pressures at the control
// microphones (Matlab/ControlMethods/predictPressure.m:12-17: p[:, m] = sum_l filter(rir[:, l, m], 1, feed[:, l])),
// acoustic contrast 10 log10(|p_bright|_F^2 / |p_dark|_F^2) and the normalised mean-square error between the target
// pressure and the bright-zone pressure (Matlab/main.m:120-130).  One pass over the loudspeaker feeds: the pressures
// are never stored, only their energies.
#include <vector>

#include "engine.cuh"

namespace apv {

namespace {

constexpr int MT = 256;   // samples per CTA

// grid (chunks, M, 2): z = 0 bright zone (+ target pressure), z = 1 dark zone.  part[chunk][m][4]:
// [0] sum p_bright^2  [1] sum (p_target - p_bright)^2  [2] sum p_target^2  [3] sum p_dark^2
__global__ void __launch_bounds__(MT) pressure_energy_kernel(const double* __restrict__ feeds, const double* __restrict__ sig,
                                                             const double* __restrict__ rirT, const double* __restrict__ rirTT,
                                                             double* __restrict__ part, int T, int K, int L, int M,
                                                             int bright) {
  extern __shared__ double sm[];
  double* hs = sm;              // reversed impulse response [K]
  double* xs = sm + K;          // feed segment [K - 1 + MT]
  __shared__ double red[40];
  const int chunk = blockIdx.x, m = blockIdx.y, which = blockIdx.z;
  const int zone = which == 0 ? bright : 1 - bright;
  const int t0 = chunk * MT, t = t0 + threadIdx.x;
  double p = 0.0;
  for (int l = 0; l < L; ++l) {
    const double* rir = rirT + (((size_t)zone * M + m) * L + l) * K;
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += MT) hs[i] = rir[K - 1 - i];
    for (int i = threadIdx.x; i < K - 1 + MT; i += MT) {
      const int ts = t0 - (K - 1) + i;
      xs[i] = (ts >= 0 && ts < T) ? feeds[(size_t)ts * L + l] : 0.0;
    }
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
    int k = 0;
    for (; k + 1 < K; k += 2) {
      a0 = fma(hs[k], xs[threadIdx.x + k], a0);
      a1 = fma(hs[k + 1], xs[threadIdx.x + k + 1], a1);
    }
    if (k < K) a0 = fma(hs[k], xs[threadIdx.x + k], a0);
    p += a0 + a1;
  }
  double pt = 0.0;
  if (which == 0) {             // target pressure: programme signal through the delayed reference RIR
    const double* rir = rirTT + ((size_t)zone * M + m) * K;
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += MT) hs[i] = rir[K - 1 - i];
    for (int i = threadIdx.x; i < K - 1 + MT; i += MT) {
      const int ts = t0 - (K - 1) + i;
      xs[i] = (ts >= 0 && ts < T) ? sig[ts] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) pt = fma(hs[k], xs[threadIdx.x + k], pt);
  }
  const bool live = t < T;
  double* P = part + ((size_t)chunk * M + m) * 4;
  if (which == 0) {
    const double e0 = block_sum(live ? p * p : 0.0, red);
    const double e1 = block_sum(live ? (pt - p) * (pt - p) : 0.0, red);
    const double e2 = block_sum(live ? pt * pt : 0.0, red);
    if (threadIdx.x == 0) { P[0] = e0; P[1] = e1; P[2] = e2; }
  } else {
    const double e3 = block_sum(live ? p * p : 0.0, red);
    if (threadIdx.x == 0) P[3] = e3;
  }
}

}  // namespace

// feeds: (T, L) host, signal: (T) host.  out[0] = acoustic contrast [dB], out[1] = mean over microphones of the
// normalised squared error (NMSE, linear), out[2] = 10 log10(NMSE) [dB].
int eval_zone(Handle& h, int zone, int T, const double* feeds, const double* signal, double* out3) {
  const Dims& D = h.D;
  if (T < 1 || zone < 0 || zone > 1) return EINVAL_;
  double *d_f = nullptr, *d_s = nullptr, *d_p = nullptr;
  const int chunks = ceil_div(T, MT);
  APV_CUDA_TRY(cudaMalloc((void**)&d_f, (size_t)T * D.L * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&d_s, (size_t)T * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&d_p, (size_t)chunks * D.M * 4 * sizeof(double)));
  APV_CUDA_TRY(cudaMemcpyAsync(d_f, feeds, (size_t)T * D.L * sizeof(double), cudaMemcpyHostToDevice, h.st));
  APV_CUDA_TRY(cudaMemcpyAsync(d_s, signal, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, h.st));
  const size_t sm = (size_t)(2 * D.K - 1 + MT) * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (sm > 48 * 1024 && sm > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(pressure_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    configured = sm;
  }
  pressure_energy_kernel<<<dim3(chunks, D.M, 2), MT, sm, h.st>>>(d_f, d_s, h.rirT, h.rirTT, d_p, T, D.K, D.L, D.M, zone);
  APV_CUDA_TRY(cudaGetLastError());
  std::vector<double> part((size_t)chunks * D.M * 4);
  APV_CUDA_TRY(cudaMemcpyAsync(part.data(), d_p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h.st));
  APV_CUDA_TRY(cudaStreamSynchronize(h.st));
  cudaFree(d_f); cudaFree(d_s); cudaFree(d_p);
  double eb = 0.0, ed = 0.0, nmse = 0.0;
  for (int m = 0; m < D.M; ++m) {
    double b = 0.0, e = 0.0, t = 0.0, dk = 0.0;
    for (int c = 0; c < chunks; ++c) {       // fixed order: deterministic
      const double* P = part.data() + ((size_t)c * D.M + m) * 4;
      b += P[0]; e += P[1]; t += P[2]; dk += P[3];
    }
    eb += b; ed += dk;
    nmse += e / t;
  }
  nmse /= D.M;
  out3[0] = 10.0 * log10(eb / ed);
  out3[1] = nmse;
  out3[2] = 10.0 * log10(nmse);
  return OK;
}

}  // namespace apv
