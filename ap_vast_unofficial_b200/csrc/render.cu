// S6  vast_sweep  -- VAST filter sum  w[v] = sum_{u<=v} (u_u^T r)/(lambda_u + mu) u_u
//                   (reference calculate_filter_spectra rank loop, Python/apvast.py:406-414), for one or
//                   several mu from ONE joint diagonalisation (U is read once per mu list entry).
// S7  render_ola  -- loudspeaker FIR rendering.  The reference multiplies rfft(win*x) with
//                   rfft(w, Nb) and overlap-adds win*irfft(.) (apvast.py:417-422,428-506), i.e. a
//                   CIRCULAR convolution (mod Nb) of the windowed input block with the J-tap filter.
//                   Here it is evaluated directly in the time domain (J taps x Nb samples per
//                   (rank, loudspeaker)), fused with the window, the in-place overlap-add shift and
//                   the (V, H, L) output transpose.
#include "engine.cuh"

namespace apv {

namespace {

// c[zi][v] = U[zi][v] . r[zone]        grid (V, nz)
__global__ void __launch_bounds__(256) sweep_dot_kernel(const double* __restrict__ U, const double* __restrict__ rvec,
                                                        double* __restrict__ cbuf, int n, int V, int zone0, int zone1) {
  __shared__ double red[40];
  const int v = blockIdx.x, zi = blockIdx.y;
  const int zone = zi == 0 ? zone0 : zone1;
  const double* u = U + ((size_t)zi * V + v) * n;
  const double* r = rvec + (size_t)zone * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(u[i], r[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) cbuf[(size_t)zi * V + v] = s;
}

// W[zone][v][i] = sum_{u<=v} c_u/(lam_u + mu) U[u][i]     grid (ceil(n/256), nz)
__global__ void __launch_bounds__(256) sweep_prefix_kernel(const double* __restrict__ U, const double* __restrict__ cbuf,
                                                           const double* __restrict__ lam, double* __restrict__ W,
                                                           double mu, int n, int V, int zone0, int zone1) {
  extern __shared__ double av[];
  const int zi = blockIdx.y;
  const int zone = zi == 0 ? zone0 : zone1;
  for (int v = threadIdx.x; v < V; v += blockDim.x) av[v] = cbuf[(size_t)zi * V + v] / (lam[(size_t)zi * V + v] + mu);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int v = 0; v < V; ++v) {
    acc = acc + av[v] * U[((size_t)zi * V + v) * n + i];
    W[((size_t)zone * V + v) * n + i] = acc;
  }
}

// S7 for the controlled streams.  grid (L, V, nz).  smem: xe[Nb + J - 1] + tr[J].
__global__ void __launch_bounds__(256) render_kernel(const double* __restrict__ xin, const double* __restrict__ W,
                                                     const double* __restrict__ win, double* __restrict__ G,
                                                     double* __restrict__ out, Dims D, int zone0, int zone1) {
  extern __shared__ double sm[];
  const int Nb = D.Nb, J = D.J, H = D.H;
  double* xe = sm;               // xe[t] = xw[(t - (J-1)) mod Nb]
  double* tr = sm + Nb + J - 1;  // tr[j'] = w[J-1-j']
  const int l = blockIdx.x, v = blockIdx.y, zone = blockIdx.z == 0 ? zone0 : zone1;
  const double* x = xin + (size_t)zone * D.LX + (D.LX - Nb);
  const double* w = W + ((size_t)zone * D.V + v) * D.n + (size_t)l * J;
  for (int t = threadIdx.x; t < Nb + J - 1; t += blockDim.x) {
    int s = t - (J - 1);
    s %= Nb;
    if (s < 0) s += Nb;
    xe[t] = win[s] * x[s];
  }
  for (int j = threadIdx.x; j < J; j += blockDim.x) tr[j] = w[J - 1 - j];
  __syncthreads();
  double* g = G + (((size_t)zone * D.V + v) * D.L + l) * Nb;
  double* o = out + ((size_t)zone * D.V + v) * H * D.L + l;
  // thread i < H owns the residue class {i, i+H, ...} of the overlap buffer
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    for (int j = i; j < Nb; j += H) {
      double a0 = 0.0, a1 = 0.0;
      const double* xp = xe + j;
      int k = 0;
      for (; k + 1 < J; k += 2) {
        a0 = fma(tr[k], xp[k], a0);
        a1 = fma(tr[k + 1], xp[k + 1], a1);
      }
      if (k < J) a0 = fma(tr[k], xp[k], a0);
      const double val = ((j + H < Nb) ? g[j + H] : 0.0) + win[j] * (a0 + a1);
      g[j] = val;
      if (j == i) o[(size_t)i * D.L] = val;
    }
  }
}

// Target streams: filter_target is a unit impulse (apvast.py:389-390), so the frame is a circular delay of the
// windowed input on one loudspeaker.  grid (2 signals).
__global__ void __launch_bounds__(256) render_target_kernel(const double* __restrict__ xin, const double* __restrict__ win,
                                                            double* __restrict__ Gt, double* __restrict__ out_t, Dims D,
                                                            int tap) {
  const int Nb = D.Nb, H = D.H;
  const int X = blockIdx.x;
  const double* x = xin + (size_t)X * D.LX + (D.LX - Nb);
  double* g = Gt + (size_t)X * Nb;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < Nb; j += H) {
      int s = (j - tap) % Nb;
      if (s < 0) s += Nb;
      const double val = ((j + H < Nb) ? g[j + H] : 0.0) + win[j] * (win[s] * x[s]);
      g[j] = val;
      if (j == i) out_t[(size_t)X * H + i] = val;
    }
}

}  // namespace

int stage_sweep(Handle& h, double mu, double* W_out) {
  const Dims& D = h.D;
  if (h.nz == 0) return OK;
  double* cbuf = h.jd.colbuf;     // free after the tridiagonalisation
  sweep_dot_kernel<<<dim3(D.V, h.nz), 256, 0, h.st>>>(h.jd.Zt, h.rvec, cbuf, D.n, D.V, h.zones[0], h.zones[1]);
  sweep_prefix_kernel<<<dim3(ceil_div(D.n, 256), h.nz), 256, D.V * sizeof(double), h.st>>>(
      h.jd.Zt, cbuf, h.jd.lam, W_out, mu, D.n, D.V, h.zones[0], h.zones[1]);
  h.launches += 2;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int stage_render(Handle& h) {
  const Dims& D = h.D;
  if (h.nz > 0) {
    const size_t sm = (size_t)(D.Nb + 2 * D.J - 1) * sizeof(double);
    static thread_local size_t configured = 0;
    if (sm > 48 * 1024 && sm > configured) {
      APV_CUDA_TRY(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      configured = sm;
    }
    render_kernel<<<dim3(D.L, D.V, h.nz), 256, sm, h.st>>>(h.xin, h.W, h.win, h.G, h.d_out, D, h.zones[0], h.zones[1]);
    h.launches += 1;
  }
  const int tgt = D.J * D.refA + D.d;      // reference uses reference_index_A for both targets (:418,422)
  render_target_kernel<<<2, 256, 0, h.st>>>(h.xin, h.win, h.Gt, h.d_out_t, D, tgt % D.J);
  h.launches += 1;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

}  // namespace apv
