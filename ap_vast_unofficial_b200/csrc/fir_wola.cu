// S1  rir_conv      streaming FIR of the new hop with every loudspeaker->microphone RIR
//                   (reference update_loudspeaker_response_buffers, Python/apvast.py:167-194:
//                    4LM + 2M scipy.signal.lfilter calls with carried state)
// S2  wola_weight   windowed FFT -> perceptual weighting -> IFFT -> window -> overlap-add ->
//                   statistics-buffer append for the target signals (apvast.py:197-235)
// S2b masking_gain  van de Par spectral-integration masking gain per microphone
//                   (apvast.py:313-327; arithmetic of Matlab/ControlMethods/perceptualModel.m:118-139,177-190)
// S3  wola_weight   same WOLA for the 4LM loudspeaker responses (apvast.py:237-311)
//
// Device layout: every buffer is channel-major and time-contiguous ([path][mic][src][t]); shifts
// ("buf <- [buf[H:]; new]") are done in place by giving each thread one residue class mod H and
// walking it in increasing order, so no ping-pong copies are needed.
#include "engine.cuh"

namespace apv {

namespace {

// ------------------------------------------------------------------------------------------
// Stockham mixed-radix FFT in shared memory, whole CTA cooperates.  tw[m] = exp(-2 pi i m / N).
// Returns the buffer holding the result (a or b).  Unnormalised in both directions.
__device__ double2* block_fft(double2* a, double2* b, const double2* __restrict__ tw, int N, int nrad,
                              const int* __restrict__ rad, bool inverse) {
  int Ns = 1;
  for (int s = 0; s < nrad; ++s) {
    const int r = rad[s];
    const int NsR = Ns * r;
    const int step = N / NsR;
    const int span = N / r;
    for (int o = threadIdx.x; o < N; o += blockDim.x) {
      const int k = o % Ns;
      const int u = (o / Ns) % r;
      const int q = o / NsR;
      const int j = q * Ns + k;
      const int base = (k + u * Ns) * step;   // < N
      double sr = 0.0, si = 0.0;
      int e = 0;
      for (int t = 0; t < r; ++t) {
        const double2 x = a[j + t * span];
        double2 w = __ldg(tw + e);
        if (inverse) w.y = -w.y;
        sr += x.x * w.x - x.y * w.y;
        si += x.x * w.y + x.y * w.x;
        e += base;
        if (e >= N) e -= N;
      }
      b[o] = make_double2(sr, si);
    }
    __syncthreads();
    double2* t = a; a = b; b = t;
    Ns = NsR;
  }
  return a;
}

struct FftPlan {
  int nrad;
  int rad[32];
};

// ------------------------------------------------------------------------------------------
// xin[X][LX] <- [xin[X][H:], new]   (also serves update_input_blocks, apvast.py:424-426)
// and xw[X][s] = win[s] * x_X[s], the windowed input block S7 filters (apvast.py:430-431): a snapshot per block, so
// that rendering block t does not depend on xin once S1 of block t+1 has shifted it.
__global__ void input_shift_kernel(double* __restrict__ xin, const double* __restrict__ inA,
                                   const double* __restrict__ inB, const double* __restrict__ win,
                                   double* __restrict__ xw, int LX, int H, int Nb) {
  double* x = xin + (size_t)blockIdx.x * LX;
  const double* in = blockIdx.x == 0 ? inA : inB;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < LX; j += H) x[j] = (j + H < LX) ? x[j + H] : in[j + H - LX];
  __syncthreads();
  const double* xb = x + (LX - Nb);
  double* o = xw + (size_t)blockIdx.x * Nb;
  for (int s = threadIdx.x; s < Nb; s += blockDim.x) o[s] = win[s] * xb[s];
}

// ------------------------------------------------------------------------------------------
// S1.  grid (L, M, 6): z<4 path p=2X+Y (signal X into zone Y), z=4,5 targets A->A, B->B (blockIdx.x==0).
__global__ void __launch_bounds__(256) fir_kernel(const double* __restrict__ xin, const double* __restrict__ rirT,
                                                  const double* __restrict__ rirTT, double* __restrict__ Q,
                                                  double* __restrict__ QT, Dims D) {
  extern __shared__ double sm[];
  const int K = D.K, H = D.H, Nb = D.Nb;
  double* hs = sm;            // reversed impulse response: hs[k] = rir[K-1-k]
  double* xs = sm + K;        // xe[0 .. K-1+H)
  const int l = blockIdx.x, m = blockIdx.y, z = blockIdx.z;
  int X;
  const double* rir;
  double* q;
  if (z < 4) {
    X = z >> 1;
    const int Y = z & 1;
    rir = rirT + (((size_t)Y * D.M + m) * D.L + l) * K;
    q = Q + (((size_t)z * D.M + m) * D.L + l) * Nb;
  } else {
    if (l != 0) return;
    X = z - 4;
    rir = rirTT + ((size_t)X * D.M + m) * K;
    q = QT + ((size_t)X * D.M + m) * Nb;
  }
  const double* xe = xin + (size_t)X * D.LX + (D.LX - (K - 1 + H));
  for (int i = threadIdx.x; i < K; i += blockDim.x) hs[i] = rir[K - 1 - i];
  for (int i = threadIdx.x; i < K - 1 + H; i += blockDim.x) xs[i] = xe[i];
  __syncthreads();
  // y[h] = sum_k rir[k] xe[K-1+h-k] = sum_j hs[j] xs[h+j]
  double* ys = xs + (K - 1 + H);   // new samples, shared so that any thread can append them
  for (int i0 = threadIdx.x; i0 < H; i0 += 4 * blockDim.x) {
    double acc[4] = {0, 0, 0, 0};
    int hh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) hh[u] = min(i0 + u * (int)blockDim.x, H - 1);
    for (int j = 0; j < K; ++j) {
      const double c = hs[j];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(c, xs[hh[u] + j], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < H) ys[i] = acc[u];
    }
  }
  __syncthreads();
  // q <- [q[H:], y]: thread i owns the residue class {i, i+H, ...} and walks it upwards (in place)
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < Nb; j += H) q[j] = (j + H < Nb) ? q[j + H] : ys[j + H - Nb];
}

// ------------------------------------------------------------------------------------------
// WOLA core shared by S2 and S3: given the (already weighted) spectrum in `spec`, inverse FFT,
// window, overlap-add into `ola` (Nb) and append the first H samples to `stats` (N).
__device__ void wola_tail(double2* spec, double2* other, const double2* tw, const double* __restrict__ win,
                          const FftPlan& pl, double* __restrict__ ola, double* __restrict__ stats, int Nb, int H,
                          int N, bool zero_frame) {
  double2* y = spec;
  if (!zero_frame) y = block_fft(spec, other, tw, Nb, pl.nrad, pl.rad, true);
  const double inv = 1.0 / Nb;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < Nb; j += H) {
      const double fr = zero_frame ? 0.0 : win[j] * (y[j].x * inv);
      ola[j] = ((j + H < Nb) ? ola[j + H] : 0.0) + fr;
    }
  __syncthreads();      // ola[0..H) complete (global writes of this CTA are visible after the barrier)
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < N; j += H) stats[j] = (j + H < N) ? stats[j + H] : ola[j + H - N];
}

// Hermitian extension of a real gain curve g[0..F) to bin f in [0, Nb)
__device__ __forceinline__ double gain_at(const double* __restrict__ g, int f, int Nb) {
  return g[f <= Nb / 2 ? f : Nb - f];
}

// S2 + S2b.  grid (M, 2 zones).  mode: 0 = W==1 ; 1 = on-device masking model ; 2 = W given (Wg is input).
// only_frame != 0: write the windowed target frame (what the reference hands to model.gain) and return.
__global__ void __launch_bounds__(256) wola_target_kernel(const double* __restrict__ QT, double* __restrict__ OT,
                                                          double* __restrict__ ST, double* __restrict__ Wg,
                                                          double* __restrict__ tframe, const double* __restrict__ win,
                                                          const double2* __restrict__ tw, FftPlan pl,
                                                          const double* __restrict__ G2, int nchan, double Cs,
                                                          double Ca, double Leff, int normalize, int mode,
                                                          int only_frame, Dims D) {
  extern __shared__ __align__(16) double sm[];
  const int Nb = D.Nb, F = D.F;
  double2* a = reinterpret_cast<double2*>(sm);
  double2* b = a + Nb;
  double* red = reinterpret_cast<double*>(b + Nb);   // 40 doubles
  double* pc = red + 40;                             // nchan doubles
  const int m = blockIdx.x, X = blockIdx.y;
  const double* q = QT + ((size_t)X * D.M + m) * Nb;
  double* wg = Wg + ((size_t)X * D.M + m) * F;
  if (only_frame) {
    double* tf = tframe + ((size_t)X * D.M + m) * Nb;
    for (int i = threadIdx.x; i < Nb; i += blockDim.x) tf[i] = win[i] * q[i];
    return;
  }
  for (int i = threadIdx.x; i < Nb; i += blockDim.x) a[i] = make_double2(win[i] * q[i], 0.0);
  __syncthreads();
  double2* s = block_fft(a, b, tw, Nb, pl.nrad, pl.rad, false);
  double2* o = (s == a) ? b : a;
  if (mode == 1) {
    // masker power per auditory channel: p_c = sum_f G2[c][f] * (2/Nb^2) |S(f)|^2
    const double sc = 2.0 / ((double)Nb * (double)Nb);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = w; c < nchan; c += nw) {
      double acc = 0.0;
      const double* g2 = G2 + (size_t)c * F;
      for (int f = lane; f < F; f += 32) {
        const double2 v = s[f];
        acc += g2[f] * (sc * (v.x * v.x + v.y * v.y));
      }
      acc = warp_sum(acc);
      if (lane == 0) pc[c] = acc;
    }
    __syncthreads();
    double nrm = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      double acc = 0.0;
      for (int c = 0; c < nchan; ++c) acc += G2[(size_t)c * F + f] / (pc[c] + Ca);
      const double g = sqrt(Cs * Leff * acc);
      wg[f] = g;
      nrm += (normalize == 2 && f > 0 && f < F - 1) ? 2.0 * g * g : g * g;     // 2: norm of the mirrored Nb-bin curve
    }
    nrm = block_sum(nrm, red);
    if (normalize) {
      const double inv = 1.0 / sqrt(nrm);
      for (int f = threadIdx.x; f < F; f += blockDim.x) wg[f] *= inv;
    }
    __syncthreads();
  }
  if (mode != 0)
    for (int f = threadIdx.x; f < Nb; f += blockDim.x) {
      const double g = gain_at(wg, f, Nb);
      s[f].x *= g;
      s[f].y *= g;
    }
  __syncthreads();
  wola_tail(s, o, tw, win, pl, OT + ((size_t)X * D.M + m) * Nb, ST + ((size_t)X * D.M + m) * D.N, Nb, D.H, D.N,
            false);
}

// S3.  grid (L, M, 4 paths).
__global__ void __launch_bounds__(256) wola_resp_kernel(const double* __restrict__ Q, double* __restrict__ O,
                                                        double* __restrict__ S, const double* __restrict__ Wg,
                                                        const double* __restrict__ win,
                                                        const double2* __restrict__ tw, FftPlan pl, int mode,
                                                        Dims D) {
  extern __shared__ __align__(16) double sm[];
  const int Nb = D.Nb;
  double2* a = reinterpret_cast<double2*>(sm);
  double2* b = a + Nb;
  const int l = blockIdx.x, m = blockIdx.y, p = blockIdx.z;
  const int X = p >> 1, Y = p & 1;
  const size_t ch = ((size_t)p * D.M + m) * D.L + l;
  const bool off = (X == 0) ? !D.runA : !D.runB;      // spectra stay zero (apvast.py:239-255)
  double2* s = a;
  double2* o = b;
  if (!off) {
    const double* q = Q + ch * Nb;
    for (int i = threadIdx.x; i < Nb; i += blockDim.x) a[i] = make_double2(win[i] * q[i], 0.0);
    __syncthreads();
    s = block_fft(a, b, tw, Nb, pl.nrad, pl.rad, false);
    o = (s == a) ? b : a;
    if (mode != 0) {
      const double* wg = Wg + ((size_t)Y * D.M + m) * D.F;   // weighting of the zone of the mic (:259-262)
      for (int f = threadIdx.x; f < Nb; f += blockDim.x) {
        const double g = gain_at(wg, f, Nb);
        s[f].x *= g;
        s[f].y *= g;
      }
    }
    __syncthreads();
  }
  wola_tail(s, o, tw, win, pl, O + ch * Nb, S + ch * D.N, Nb, D.H, D.N, off);
}

__global__ void fft_util_kernel(const double2* in, double2* out, const double2* tw, FftPlan pl, int N, int inverse) {
  extern __shared__ __align__(16) double sm[];
  double2* a = reinterpret_cast<double2*>(sm);
  double2* b = a + N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) a[i] = in[i];
  __syncthreads();
  double2* s = block_fft(a, b, tw, N, pl.nrad, pl.rad, inverse != 0);
  for (int i = threadIdx.x; i < N; i += blockDim.x) out[i] = s[i];
}

FftPlan make_plan(const Handle& h) {
  FftPlan p;
  p.nrad = h.nrad;
  for (int i = 0; i < 32; ++i) p.rad[i] = i < h.nrad ? h.rad[i] : 1;
  return p;
}

template <typename Kern>
int ensure_smem(Kern k, size_t bytes) {
  if (bytes > 48 * 1024) APV_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return OK;
}

}  // namespace

// Factor n into radices (4 first, then 2, 3, 5, then remaining primes).
int fft_plan(int n, int* rad, int* nrad) {
  int c = 0, m = n;
  while (m % 4 == 0 && c < 32) { rad[c++] = 4; m /= 4; }
  for (int p = 2; m > 1 && c < 32;) {
    if (m % p == 0) { rad[c++] = p; m /= p; }
    else { p += (p == 2) ? 1 : 2; if ((long long)p * p > m) p = m; }
  }
  if (m != 1) return EINVAL_;
  *nrad = c;
  return OK;
}

int stage_fir(Handle& h, const double* d_inA, const double* d_inB) {
  const Dims& D = h.D;
  input_shift_kernel<<<2, 256, 0, h.st>>>(h.xin, d_inA, d_inB, h.win, h.xw, D.LX, D.H, D.Nb);
  size_t sm = (size_t)(2 * D.K - 1 + 2 * D.H) * sizeof(double);
  APV_TRY(ensure_smem(fir_kernel, sm));
  fir_kernel<<<dim3(D.L, D.M, 6), 256, sm, h.st>>>(h.xin, h.rirT, h.rirTT, h.Q, h.QT, D);
  h.launches += 2;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int stage_targets(Handle& h, bool only_frame) {
  const Dims& D = h.D;
  size_t sm = (size_t)D.Nb * 2 * sizeof(double2) + (40 + (h.nchan > 0 ? h.nchan : 1)) * sizeof(double);
  APV_TRY(ensure_smem(wola_target_kernel, sm));
  wola_target_kernel<<<dim3(D.M, 2), 256, sm, h.st>>>(h.QT, h.OT, h.ST, h.Wg, h.tframe, h.win, h.tw, make_plan(h),
                                                      h.G2, h.nchan, h.Cs, h.Ca, h.Leff, h.cfg.normalize_gains,
                                                      h.cfg.perceptual == 3 ? 1 : h.cfg.perceptual, only_frame ? 1 : 0, D);
  h.launches += 1;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int stage_weighted(Handle& h) {
  const Dims& D = h.D;
  size_t sm = (size_t)D.Nb * 2 * sizeof(double2);
  APV_TRY(ensure_smem(wola_resp_kernel, sm));
  wola_resp_kernel<<<dim3(D.L, D.M, 4), 256, sm, h.st>>>(h.Q, h.O, h.S, h.Wg, h.win, h.tw, make_plan(h),
                                                         h.cfg.perceptual, D);
  h.launches += 1;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int fft_util(int n, int inverse, const double* in_ri, double* out_ri) {
  if (n < 1) return EINVAL_;
  FftPlan pl;
  for (int i = 0; i < 32; ++i) pl.rad[i] = 1;
  APV_TRY(fft_plan(n, pl.rad, &pl.nrad));
  double2 *d_in = nullptr, *d_out = nullptr, *d_tw = nullptr;
  double2* tw = (double2*)malloc(sizeof(double2) * n);
  for (int k = 0; k < n; ++k) {
    long double ang = -2.0L * 3.141592653589793238462643383279502884L * k / n;
    tw[k] = make_double2((double)cosl(ang), (double)sinl(ang));
  }
  APV_CUDA_TRY(cudaMalloc(&d_in, sizeof(double2) * n));
  APV_CUDA_TRY(cudaMalloc(&d_out, sizeof(double2) * n));
  APV_CUDA_TRY(cudaMalloc(&d_tw, sizeof(double2) * n));
  APV_CUDA_TRY(cudaMemcpy(d_in, in_ri, sizeof(double2) * n, cudaMemcpyHostToDevice));
  APV_CUDA_TRY(cudaMemcpy(d_tw, tw, sizeof(double2) * n, cudaMemcpyHostToDevice));
  free(tw);
  size_t sm = (size_t)n * 2 * sizeof(double2);
  APV_TRY(ensure_smem(fft_util_kernel, sm));
  fft_util_kernel<<<1, 256, sm>>>(d_in, d_out, d_tw, pl, n, inverse);
  APV_CUDA_TRY(cudaGetLastError());
  APV_CUDA_TRY(cudaMemcpy(out_ri, d_out, sizeof(double2) * n, cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_tw);
  return OK;
}

}  // namespace apv
