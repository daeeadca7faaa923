// S6  vast_sweep  -- VAST filter sum  w[v] = sum_{u<=v} (u_u^T r)/(lambda_u + mu) u_u
//                   (reference calculate_filter_spectra rank loop, Python/apvast.py:406-414), for one or
//                   several mu from ONE joint diagonalisation (U is read once per mu list entry).
// S7  render_ola  -- loudspeaker FIR rendering.  The reference multiplies rfft(win*x) with
//                   rfft(w, Nb) and overlap-adds win*irfft(.) (apvast.py:417-422,428-506), i.e. a
//                   CIRCULAR convolution (mod Nb) of the windowed input block with the J-tap filter.
//                   Here it is evaluated directly in the time domain (J taps x Nb samples per
//                   (rank, loudspeaker)), fused with the window, the in-place overlap-add shift and
//                   the (V, H, L) output transpose.
#include <algorithm>

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

// The sweep for a LIST of mu in one launch (BASELINE cfg-4): W[k][zone][v][i] = sum_{u<=v} c_u / (lam_u + mu_k) U[u][i].
// grid (ceil(n / 64), nz, n_mu), 256 threads = 64 columns i x 4 rank segments: a thread first sums its segment, the
// segment totals are exchanged through shared memory, then the segment is walked again with its offset and written
// (U is read twice, W written once: ~3 V n 8 bytes per mu and zone, all coalesced 512-byte rows).
__global__ void __launch_bounds__(256) sweep_multi_kernel(const double* __restrict__ U, const double* __restrict__ cbuf,
                                                          const double* __restrict__ lam, const double* __restrict__ mu,
                                                          double* __restrict__ W, int n, int V, int zone0, int zone1) {
  extern __shared__ double av[];                  // a[v] = c_v / (lam_v + mu), V doubles, then 4 x 64 segment totals
  double* tot = av + V;
  const int zi = blockIdx.y, k = blockIdx.z;
  const int zone = zi == 0 ? zone0 : zone1;
  const double m = mu[k];
  for (int v = threadIdx.x; v < V; v += blockDim.x) av[v] = cbuf[(size_t)zi * V + v] / (lam[(size_t)zi * V + v] + m);
  __syncthreads();
  const int il = threadIdx.x & 63, seg = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + il;
  const int slen = (V + 3) / 4, v0 = seg * slen, v1 = min(V, v0 + slen);
  const double* u = U + (size_t)zi * V * n + i;
  double acc = 0.0;
  if (i < n)
    for (int v = v0; v < v1; ++v) acc = fma(av[v], u[(size_t)v * n], acc);
  tot[seg * 64 + il] = acc;
  __syncthreads();
  if (i >= n) return;
  acc = 0.0;
  for (int q = 0; q < seg; ++q) acc += tot[q * 64 + il];
  double* w = W + (((size_t)k * 2 + zone) * V) * n + i;
  for (int v = v0; v < v1; ++v) {
    acc = fma(av[v], u[(size_t)v * n], acc);
    w[(size_t)v * n] = acc;
  }
}

// Eigen-basis figures of merit of the mu x V sweep (BASELINE cfg-4): with U^T (R_D + reg I) U = I and
// U^T R_B U = Lambda the rank-v filter w = sum_{i<=v} a_i u_i, a_i = c_i / (lambda_i + mu), has
//   dark energy   w^T (R_D + reg I) w = sum a_i^2,   bright energy  w^T R_B w = sum lambda_i a_i^2,   w^T r_B = sum a_i c_i.
// out[k][zone][v][3] for mu_k.   grid (n_mu, nz); the prefix over the ranks is sequential (V terms, one thread).
__global__ void sweep_metrics_kernel(const double* __restrict__ cbuf, const double* __restrict__ lam,
                                     const double* __restrict__ mu, double* __restrict__ out, int V, int zone0, int zone1) {
  const int k = blockIdx.x, zi = blockIdx.y;
  const int zone = zi == 0 ? zone0 : zone1;
  if (threadIdx.x != 0) return;
  const double m = mu[k];
  double dark = 0.0, bright = 0.0, cross = 0.0;
  double* o = out + (((size_t)k * 2 + zone) * V) * 3;
  for (int v = 0; v < V; ++v) {
    const double c = cbuf[(size_t)zi * V + v], l = lam[(size_t)zi * V + v];
    const double a = c / (l + m);
    dark = fma(a, a, dark);
    bright = fma(l * a, a, bright);
    cross = fma(a, c, cross);
    o[3 * v] = dark; o[3 * v + 1] = bright; o[3 * v + 2] = cross;
  }
}

// S7 for the controlled streams.  grid (H / RT, V, nz), 256 threads.  A CTA renders a tile of RT hop positions (and
// the Nb/H - 1 overlap positions behind each of them) of ALL loudspeakers of one (zone, rank): lane = hop position, so
// the overlap buffer is read and written with unit stride and the tap loop reads the filter by broadcast; every warp
// works on two loudspeakers at a time (one shared-memory load per FMA instead of two).  The (RT, L) output tile is
// staged in shared memory and leaves as one contiguous run of double2 (the (V, H, L) layout of the reference's output
// lists).  Only the first min(J, Nb) taps act: rfft(w, Nb) crops the filter (apvast.py:417-420).
constexpr int RT = 64;       // hop positions per CTA

struct RenderPlan {
  int LC;       // loudspeakers per shared-memory pass
  int Je;       // taps that act = min(J, Nb)
  int ncls;     // residue classes = ceil(Nb / H)
  size_t smem;
};

__global__ void __launch_bounds__(256) render_kernel(const double* __restrict__ xw, const double* __restrict__ W,
                                                     const double* __restrict__ win, double* __restrict__ G,
                                                     double* __restrict__ out, Dims D, int zone0, int zone1, int LC,
                                                     int Je, int ncls) {
  extern __shared__ __align__(16) double sm[];
  const int Nb = D.Nb, J = D.J, H = D.H, L = D.L;
  const int XS = (RT + Je + 1) & ~1;              // padded (even) length of one class window
  double* tr = sm;                                // [LC][Je]   tr[l][k'] = w[l][Je-1-k']
  double* xe = tr + (((size_t)LC * Je + 1) & ~(size_t)1);   // even offsets: the output tile is read as double2              // [ncls][XS] xe[c][t] = xw[(h0 + c H + t - (Je-1)) mod Nb]
  double* ot = xe + (size_t)ncls * XS;            // [RT][L]    output tile
  const int h0 = blockIdx.x * RT, v = blockIdx.y, zone = blockIdx.z == 0 ? zone0 : zone1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* x = xw + (size_t)zone * Nb;
  for (int e = tid; e < ncls * XS; e += 256) {
    const int c = e / XS, t = e - c * XS;
    int s = (h0 + c * H + t - (Je - 1)) % Nb;
    if (s < 0) s += Nb;
    xe[e] = (t < RT + Je - 1) ? x[s] : 0.0;
  }
  const double* wz = W + ((size_t)zone * D.V + v) * D.n;
  double* gz = G + ((size_t)zone * D.V + v) * L * Nb;
  for (int l0 = 0; l0 < L; l0 += LC) {
    const int lc = min(LC, L - l0);
    __syncthreads();
    for (int e = tid; e < lc * Je; e += 256) {
      const int l = e / Je, k = e - l * Je;
      tr[e] = wz[(size_t)(l0 + l) * J + (Je - 1 - k)];
    }
    __syncthreads();
    for (int lp = 2 * warp; lp < lc; lp += 16) {          // loudspeaker pair (lp, lp + 1)
      const bool two = lp + 1 < lc;
      const double* t0 = tr + (size_t)lp * Je;
      const double* t1 = tr + (size_t)(two ? lp + 1 : lp) * Je;
      for (int c = 0; c < ncls; ++c) {
        const double* xc = xe + (size_t)c * XS + lane;
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;     // [loudspeaker][hop position lane / lane + 32]
        double b00 = 0.0, b01 = 0.0, b10 = 0.0, b11 = 0.0;     // odd taps
        int k = 0;
        for (; k + 1 < Je; k += 2) {
          const double w0 = t0[k], w0b = t0[k + 1], w1 = t1[k], w1b = t1[k + 1];
          const double x0 = xc[k], x0b = xc[k + 1], x1 = xc[k + 32], x1b = xc[k + 33];
          a00 = fma(w0, x0, a00); a01 = fma(w0, x1, a01);
          a10 = fma(w1, x0, a10); a11 = fma(w1, x1, a11);
          b00 = fma(w0b, x0b, b00); b01 = fma(w0b, x1b, b01);
          b10 = fma(w1b, x0b, b10); b11 = fma(w1b, x1b, b11);
        }
        if (k < Je) {
          const double w0 = t0[k], w1 = t1[k], x0 = xc[k], x1 = xc[k + 32];
          a00 = fma(w0, x0, a00); a01 = fma(w0, x1, a01);
          a10 = fma(w1, x0, a10); a11 = fma(w1, x1, a11);
        }
        const double y[2][2] = {{a00 + b00, a01 + b01}, {a10 + b10, a11 + b11}};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !two) break;
          double* g = gz + (size_t)(l0 + lp + q) * Nb;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int hl = lane + 32 * u, h = h0 + hl;
            const int j = h + c * H;
            if (h < H && j < Nb) {
              // in-place overlap-add shift: class c of position h is read (g[j + H]) before class c + 1 overwrites it
              const double val = ((j + H < Nb) ? g[j + H] : 0.0) + win[j] * y[q][u];
              g[j] = val;
              if (c == 0) ot[(size_t)hl * L + l0 + lp + q] = val;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int rows = min(RT, H - h0);
  double* o = out + ((size_t)zone * D.V + v) * H * L + (size_t)h0 * L;
  const int cnt = rows * L;
  if ((cnt & 1) == 0 && ((reinterpret_cast<size_t>(o) & 15) == 0)) {
    double2* o2 = reinterpret_cast<double2*>(o);
    const double2* s2 = reinterpret_cast<const double2*>(ot);
    for (int e = tid; e < cnt / 2; e += 256) o2[e] = s2[e];
  } else {
    for (int e = tid; e < cnt; e += 256) o[e] = ot[e];
  }
}

// Target streams: filter_target is a unit impulse (apvast.py:389-390), so the frame is a circular delay of the
// windowed input on one loudspeaker.  grid (2 signals).
__global__ void __launch_bounds__(256) render_target_kernel(const double* __restrict__ xw, const double* __restrict__ win,
                                                            double* __restrict__ Gt, double* __restrict__ out_t, Dims D,
                                                            int tap) {
  const int Nb = D.Nb, H = D.H;
  const int X = blockIdx.x;
  const double* x = xw + (size_t)X * Nb;
  double* g = Gt + (size_t)X * Nb;
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    for (int j = i; j < Nb; j += H) {
      int s = (j - tap) % Nb;
      if (s < 0) s += Nb;
      const double val = ((j + H < Nb) ? g[j + H] : 0.0) + win[j] * x[s];
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

// filters for n_mu values of mu (device array) into W_out (n_mu, 2, V, n); c = U^T r is formed once
int stage_sweep_multi(Handle& h, int n_mu, const double* d_mu, double* W_out) {
  const Dims& D = h.D;
  if (h.nz == 0) return OK;
  double* cbuf = h.jd.colbuf;
  sweep_dot_kernel<<<dim3(D.V, h.nz), 256, 0, h.st>>>(h.jd.Zt, h.rvec, cbuf, D.n, D.V, h.zones[0], h.zones[1]);
  const size_t sm = (size_t)(D.V + 256) * sizeof(double);
  static PerDevice pd_configured; size_t& configured = pd_configured.cur();
  if (sm > 48 * 1024 && sm > configured) {
    APV_CUDA_TRY(cudaFuncSetAttribute(sweep_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    configured = sm;
  }
  sweep_multi_kernel<<<dim3(ceil_div(D.n, 64), h.nz, n_mu), 256, sm, h.st>>>(h.jd.Zt, cbuf, h.jd.lam, d_mu, W_out, D.n, D.V,
                                                                             h.zones[0], h.zones[1]);
  h.launches += 2;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int stage_sweep_metrics(Handle& h, int n_mu, const double* d_mu, double* d_out) {
  const Dims& D = h.D;
  if (h.nz == 0) return OK;
  double* cbuf = h.jd.colbuf;
  sweep_dot_kernel<<<dim3(D.V, h.nz), 256, 0, h.st>>>(h.jd.Zt, h.rvec, cbuf, D.n, D.V, h.zones[0], h.zones[1]);
  sweep_metrics_kernel<<<dim3(n_mu, h.nz), 32, 0, h.st>>>(cbuf, h.jd.lam, d_mu, d_out, D.V, h.zones[0], h.zones[1]);
  h.launches += 2;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

int stage_render(Handle& h) {
  const Dims& D = h.D;
  if (h.nz > 0) {
    RenderPlan rp;
    rp.Je = std::min(D.J, D.Nb);
    rp.ncls = ceil_div(D.Nb, D.H);
    const size_t fixed = ((size_t)rp.ncls * ((RT + rp.Je + 1) & ~1) + (size_t)RT * D.L + 2) * sizeof(double);
    rp.LC = D.L;
    while (rp.LC > 2 && fixed + (size_t)rp.LC * rp.Je * sizeof(double) > 160 * 1024) rp.LC = (rp.LC + 1) / 2;
    rp.LC = std::max(2, (rp.LC + 1) & ~1);       // warps work on loudspeaker pairs
    rp.smem = fixed + (size_t)rp.LC * rp.Je * sizeof(double);
    if (rp.smem > 220 * 1024) {
      snprintf(g_err, sizeof(g_err), "render: block/filter sizes need %zu B of shared memory", rp.smem);
      return EINVAL_;
    }
    static PerDevice pd_configured; size_t& configured = pd_configured.cur();
    if (rp.smem > 48 * 1024 && rp.smem > configured) {
      APV_CUDA_TRY(cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp.smem));
      configured = rp.smem;
    }
    render_kernel<<<dim3(ceil_div(D.H, RT), D.V, h.nz), 256, rp.smem, h.st>>>(h.xw, h.W, h.win, h.G, h.d_out, D, h.zones[0],
                                                                            h.zones[1], rp.LC, rp.Je, rp.ncls);
    h.launches += 1;
  }
  const int tgt = D.J * D.refA + D.d;      // reference uses reference_index_A for both targets (:418,422)
  render_target_kernel<<<2, 256, 0, h.st>>>(h.xw, h.win, h.Gt, h.d_out_t, D, tgt % D.J);
  h.launches += 1;
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

}  // namespace apv
