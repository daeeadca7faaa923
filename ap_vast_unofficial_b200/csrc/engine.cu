// C-ABI of the AP-VAST B200 engine (include/apvast_b200.h): handle life-cycle, the per-block call that
// chains S1..S7 on one CUDA stream, state get/set, and the test / measurement utilities.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "engine.cuh"

using namespace apv;

namespace {

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(apv::g_err, sizeof(apv::g_err), fmt, ap);
  va_end(ap);
  return code;
}

template <typename T>
int dalloc(T** p, size_t count) {
  if (count == 0) count = 1;
  APV_CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  APV_CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
  return OK;
}

__global__ void fill_kernel(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

struct TensorInfo {
  double* ptr;
  size_t count;
};

TensorInfo tensor_info(const Handle& h, int id) {
  const Dims& D = h.D;
  const size_t M = D.M, L = D.L, Nb = D.Nb, N = D.N, V = D.V, n = D.n, F = D.F;
  switch (id) {
    case APV_T_W: return {h.W, 2 * V * n};
    case APV_T_LAMBDA: return {h.lam, 2 * V};
    case APV_T_U: return {h.U, 2 * V * n};
    case APV_T_R: return {h.R, 4 * n * n};          // strided copy (ldn) handled by the caller
    case APV_T_RVEC: return {h.rvec, 2 * n};
    case APV_T_WEIGHT: return {h.Wg, 2 * M * F};
    case APV_T_RESP: return {h.Q, 4 * M * L * Nb};
    case APV_T_RESP_T: return {h.QT, 2 * M * Nb};
    case APV_T_OLA: return {h.O, 4 * M * L * Nb};
    case APV_T_OLA_T: return {h.OT, 2 * M * Nb};
    case APV_T_STATS: return {h.S, 4 * M * L * N};
    case APV_T_STATS_T: return {h.ST, 2 * M * N};
    case APV_T_OUT_OLA: return {h.G, 2 * V * L * Nb};
    case APV_T_OUT_OLA_T: return {h.Gt, 2 * Nb};
    case APV_T_INPUT: return {h.xin, 2 * (size_t)D.LX};
    case APV_T_TARGET_FRAME: return {h.tframe, 2 * M * Nb};
    default: return {nullptr, 0};
  }
}

int ensure_pinned(Handle& h, size_t count) {
  if (h.h_pin_count >= count) return OK;
  if (h.h_pin) cudaFreeHost(h.h_pin);
  h.h_pin = nullptr;
  h.h_pin_count = 0;
  APV_CUDA_TRY(cudaMallocHost((void**)&h.h_pin, count * sizeof(double)));
  h.h_pin_count = count;
  return OK;
}

// copy the per-zone results of the joint diagonalisation into the zone-indexed result arrays
int publish_eig(Handle& h) {
  const Dims& D = h.D;
  for (int zi = 0; zi < h.nz; ++zi) {
    const int zone = h.zones[zi];
    APV_CUDA_TRY(cudaMemcpyAsync(h.lam + (size_t)zone * D.V, h.jd.lam + (size_t)zi * D.V, D.V * sizeof(double),
                                 cudaMemcpyDeviceToDevice, h.st));
    APV_CUDA_TRY(cudaMemcpyAsync(h.U + (size_t)zone * D.V * D.n, h.jd.Zt + (size_t)zi * D.V * D.n,
                                 (size_t)D.V * D.n * sizeof(double), cudaMemcpyDeviceToDevice, h.st));
  }
  return OK;
}

int run_jdiag(Handle& h) {
  if (h.nz == 0) return OK;
  const Dims& D = h.D;
  const size_t ms = (size_t)D.n * D.ldn;
  // zone A: bright R_A_to_A (path 0), dark R_A_to_B (path 1); zone B: bright R_B_to_B (3), dark R_B_to_A (2)
  const double* bright[2];
  const double* dark[2];
  for (int zi = 0; zi < h.nz; ++zi) {
    const int zone = h.zones[zi];
    bright[zi] = h.R + (zone == 0 ? 0 : 3) * ms;
    dark[zi] = h.R + (zone == 0 ? 1 : 2) * ms;
  }
  if (h.nz == 1) { bright[1] = bright[0]; dark[1] = dark[0]; }
  return jdiag_run(h.jd, bright, dark, D.ldn, h.cfg.loading_mode == 1 ? 0.0 : h.cfg.reg, h.st, &h.launches);
}

int check_info(Handle& h) {
  if (h.nz == 0) return OK;
  int info[8] = {0};
  APV_CUDA_TRY(cudaMemcpyAsync(info, h.jd.info, (size_t)h.nz * 4 * sizeof(int), cudaMemcpyDeviceToHost, h.st));
  APV_CUDA_TRY(cudaStreamSynchronize(h.st));
  for (int zi = 0; zi < h.nz; ++zi) {
    if (info[zi * 4] != 0)
      return fail(ENOTPD, "Matrix is not positive definite (zone %c, pivot %d)", h.zones[zi] == 0 ? 'A' : 'B',
                  info[zi * 4]);
    if (info[zi * 4 + 1] != 0)
      return fail(ENOCONV, "eigen-solver did not converge (zone %c, flags %d)", h.zones[zi] == 0 ? 'A' : 'B',
                  info[zi * 4 + 1]);
  }
  return OK;
}

// S1..S7 with the inputs already on the device.  `from_targets`: S1 was already run (split call).
int run_block(Handle& h, const double* d_inA, const double* d_inB, bool skip_s1, bool state_only) {
  cudaEvent_t* ev = h.ev;
  h.launches = 0;
  APV_CUDA_TRY(cudaEventRecord(ev[0], h.st));
  if (!skip_s1) APV_TRY(stage_fir(h, d_inA, d_inB));
  APV_CUDA_TRY(cudaEventRecord(ev[1], h.st));
  APV_TRY(stage_targets(h, false));
  APV_TRY(stage_weighted(h));
  APV_CUDA_TRY(cudaEventRecord(ev[2], h.st));
  if (!state_only) {
    APV_TRY(stage_stats(h));
    if (h.cfg.loading_mode == 1) APV_TRY(stage_loading(h));
    APV_CUDA_TRY(cudaEventRecord(ev[3], h.st));
    APV_TRY(run_jdiag(h));
    APV_TRY(publish_eig(h));
    APV_CUDA_TRY(cudaEventRecord(ev[4], h.st));
    APV_TRY(stage_sweep(h, h.cfg.mu, h.W));
    APV_CUDA_TRY(cudaEventRecord(ev[5], h.st));
    APV_TRY(stage_render(h));
  } else {
    for (int i = 3; i <= 5; ++i) APV_CUDA_TRY(cudaEventRecord(ev[i], h.st));
  }
  APV_CUDA_TRY(cudaEventRecord(ev[6], h.st));
  return OK;
}

int copy_in(Handle& h, const double* in_A, const double* in_B) {
  const int H = h.D.H;
  APV_TRY(ensure_pinned(h, 2 * (size_t)H));
  memcpy(h.h_pin, in_A, H * sizeof(double));
  memcpy(h.h_pin + H, in_B, H * sizeof(double));
  APV_CUDA_TRY(cudaMemcpyAsync(h.d_in, h.h_pin, 2 * (size_t)H * sizeof(double), cudaMemcpyHostToDevice, h.st));
  return OK;
}

int copy_out(Handle& h, double* out_A, double* out_B, double* out_A_t, double* out_B_t) {
  const Dims& D = h.D;
  const size_t per = (size_t)D.V * D.H * D.L;
  if (out_A && D.runA)
    APV_CUDA_TRY(cudaMemcpyAsync(out_A, h.d_out, per * sizeof(double), cudaMemcpyDeviceToHost, h.st));
  if (out_B && D.runB)
    APV_CUDA_TRY(cudaMemcpyAsync(out_B, h.d_out + per, per * sizeof(double), cudaMemcpyDeviceToHost, h.st));
  std::vector<double> t;
  if (out_A_t || out_B_t) {
    t.resize(2 * (size_t)D.H);
    APV_CUDA_TRY(cudaMemcpyAsync(t.data(), h.d_out_t, 2 * (size_t)D.H * sizeof(double), cudaMemcpyDeviceToHost, h.st));
  }
  APV_CUDA_TRY(cudaStreamSynchronize(h.st));
  for (int X = 0; X < 2; ++X) {
    double* o = X == 0 ? out_A_t : out_B_t;
    if (!o) continue;
    const int ref = (X == 1 && h.cfg.target_ref_per_zone) ? D.refB : D.refA;   // Python uses index A for both (:418,422)
    const int lt = (D.J * ref + D.d) / D.J;
    memset(o, 0, (size_t)D.H * D.L * sizeof(double));
    for (int i = 0; i < D.H; ++i) o[(size_t)i * D.L + lt] = t[(size_t)X * D.H + i];
  }
  return OK;
}

}  // namespace

extern "C" {

const char* apv_last_error(void) { return apv::g_err; }
const char* apv_version(void) { return "apvast_b200 0.1 (sm_100a)"; }

size_t apv_tensor_size(const apv_handle* h, int id) {
  if (!h) return 0;
  return tensor_info(*h, id).count;
}

int apv_create(const apv_config* cfg, const double* rir_A, const double* rir_B, const double* init_resp,
               apv_handle** out) {
  if (!cfg || !rir_A || !rir_B || !out) return fail(EINVAL_, "null argument");
  *out = nullptr;
  if (cfg->block_size <= 0 || cfg->block_size % 2 != 0) return fail(EINVAL_, "block size must be modulo 2");
  Dims D{};
  D.Nb = cfg->block_size;
  D.H = cfg->hop_size > 0 ? cfg->hop_size : D.Nb / 2;
  D.K = cfg->rir_length; D.L = cfg->n_srcs; D.M = cfg->n_mics; D.J = cfg->filter_length;
  D.N = cfg->stats_length; D.V = cfg->n_eig; D.d = cfg->modeling_delay;
  D.refA = cfg->ref_A; D.refB = cfg->ref_B; D.runA = cfg->run_A != 0; D.runB = cfg->run_B != 0;
  D.F = D.Nb / 2 + 1;
  D.n = D.L * D.J;
  D.ldn = round_up(D.n, 8);
  D.clean = cfg->toeplitz_clean != 0;
  D.P = D.N - D.J + (D.clean ? 1 : 0);
  if (D.K < 1 || D.L < 1 || D.M < 1 || D.J < 1 || D.V < 1) return fail(EINVAL_, "non-positive dimension");
  if (D.H > D.Nb) return fail(EINVAL_, "hop size larger than block size");
  if (D.H > D.N) return fail(EINVAL_, "hop size larger than statistics buffer");
  if (D.P < 1) return fail(EINVAL_, "statistics_buffer_length must exceed filter_length");
  if (D.V > D.n) return fail(EINVAL_, "number_of_eigenvectors exceeds filter_length * number_of_srcs");
  if (D.refA < 0 || D.refA >= D.L || D.refB < 0 || D.refB >= D.L) return fail(EINVAL_, "reference index out of range");
  if (D.d < 0 || D.d > D.K) return fail(EINVAL_, "modeling delay out of range");
  if (D.J * D.refA + D.d >= D.n) return fail(EINVAL_, "target tap index out of range");
  D.Ns = round_up(ceil_div(D.P, STATS_KC) * STATS_KC + D.J + 2, 8);
  D.LX = D.K - 1 + D.H > D.Nb ? D.K - 1 + D.H : D.Nb;

  apv_handle* h = new apv_handle();
  h->cfg = *cfg;
  h->cfg.hop_size = D.H;
  if (h->cfg.reg <= 0) h->cfg.reg = 1e-7;
  if (h->cfg.bright_load <= 0) h->cfg.bright_load = 1e-8;
  if (h->cfg.dark_load <= 0) h->cfg.dark_load = 5e-3;
  h->D = D;
  auto bail = [&](int code) { apv_destroy(h); return code; };
  if (cfg->device >= 0) {
    if (cudaSetDevice(cfg->device) != cudaSuccess) return bail(fail(ECUDA, "cudaSetDevice(%d) failed", cfg->device));
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return bail(fail(ECUDA, "no CUDA device: the AP-VAST B200 engine has no CPU fallback"));
  cudaGetDevice(&h->device);
#define TRYB(x) do { int _s = (x); if (_s != OK) return bail(_s); } while (0)
#define CUB(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return bail(fail(ECUDA, "%s -> %s", #x, cudaGetErrorString(_e))); } while (0)
  CUB(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  for (auto& e : h->ev) CUB(cudaEventCreate(&e));
  for (auto& e : h->ev_syrk) CUB(cudaEventCreate(&e));
  for (auto& e : h->ev_timer) CUB(cudaEventCreate(&e));
  const size_t K = D.K, L = D.L, M = D.M, Nb = D.Nb, N = D.N, V = D.V, n = D.n;
  TRYB(dalloc(&h->rirT, 2 * M * L * K));
  TRYB(dalloc(&h->rirTT, 2 * M * K));
  TRYB(dalloc(&h->win, Nb));
  TRYB(dalloc(&h->tw, Nb));
  TRYB(dalloc(&h->xin, 2 * (size_t)D.LX));
  TRYB(dalloc(&h->Q, 4 * M * L * Nb));
  TRYB(dalloc(&h->QT, 2 * M * Nb));
  TRYB(dalloc(&h->O, 4 * M * L * Nb));
  TRYB(dalloc(&h->OT, 2 * M * Nb));
  TRYB(dalloc(&h->S, 4 * M * L * N));
  TRYB(dalloc(&h->ST, 2 * M * N));
  TRYB(dalloc(&h->Sp, 4 * M * L * (size_t)D.Ns));
  TRYB(dalloc(&h->Wg, 2 * M * (size_t)D.F));
  TRYB(dalloc(&h->seed, 4 * L * L * (size_t)D.J));
  if (cfg->stats_mode != 2) TRYB(dalloc(&h->Pbuf, 16 * n * (size_t)D.ldn));
  TRYB(dalloc(&h->norms, 64));
  TRYB(dalloc(&h->pvec, 8 * n));
  TRYB(dalloc(&h->tframe, 2 * M * Nb));
  TRYB(dalloc(&h->G, 2 * V * L * Nb));
  TRYB(dalloc(&h->Gt, 2 * Nb));
  TRYB(dalloc(&h->R, 4 * n * (size_t)D.ldn));
  TRYB(dalloc(&h->rvec, 2 * n));
  TRYB(dalloc(&h->lam, 2 * V));
  TRYB(dalloc(&h->U, 2 * V * n));
  TRYB(dalloc(&h->W, 2 * V * n));
  TRYB(dalloc(&h->d_in, 2 * (size_t)D.H));
  TRYB(dalloc(&h->d_out, 2 * V * (size_t)D.H * L));
  TRYB(dalloc(&h->d_out_t, 2 * (size_t)D.H));
  h->nz = 0;
  if (D.runA) h->zones[h->nz++] = 0;
  if (D.runB) h->zones[h->nz++] = 1;
  if (h->nz == 1) h->zones[1] = h->zones[0];
  if (h->nz > 0) TRYB(jdiag_alloc(h->jd, D.n, D.V, h->nz, cfg->eig_mode));
  if (fft_plan(D.Nb, h->rad, &h->nrad) != OK) return bail(fail(EINVAL_, "cannot factor block size %d", D.Nb));

  // host-side constant tables
  std::vector<double> buf;
  {  // window sin(pi n / Nb) (apvast.py:94) and twiddles
    buf.resize(Nb);
    for (size_t i = 0; i < Nb; ++i) buf[i] = sin(M_PI / (double)Nb * (double)i);
    CUB(cudaMemcpy(h->win, buf.data(), Nb * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<double2> tw(Nb);
    for (size_t k = 0; k < Nb; ++k) {
      const long double ang = -2.0L * 3.141592653589793238462643383279502884L * (long double)k / (long double)Nb;
      tw[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    CUB(cudaMemcpy(h->tw, tw.data(), Nb * sizeof(double2), cudaMemcpyHostToDevice));
  }
  {  // RIRs (K, L, M) -> [zone][m][l][k]; delayed target RIRs (apvast.py:102-112)
    buf.assign(2 * M * L * K, 0.0);
    std::vector<double> tt(2 * M * K, 0.0);
    for (int Y = 0; Y < 2; ++Y) {
      const double* r = Y == 0 ? rir_A : rir_B;
      const int ref = Y == 0 ? D.refA : D.refB;
      for (size_t k = 0; k < K; ++k)
        for (size_t l = 0; l < L; ++l)
          for (size_t m = 0; m < M; ++m) buf[((Y * M + m) * L + l) * K + k] = r[(k * L + l) * M + m];
      for (size_t m = 0; m < M; ++m)
        for (size_t k = D.d; k < K; ++k) tt[(Y * M + m) * K + k] = r[((k - D.d) * L + ref) * M + m];
    }
    CUB(cudaMemcpy(h->rirT, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(h->rirTT, tt.data(), tt.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (init_resp) {  // 4 x (Nb, L, M) then 2 x (Nb, M)  ->  [path][m][l][t], [zone][m][t]
    buf.assign(4 * M * L * Nb, 0.0);
    for (size_t p = 0; p < 4; ++p)
      for (size_t t = 0; t < Nb; ++t)
        for (size_t l = 0; l < L; ++l)
          for (size_t m = 0; m < M; ++m)
            buf[((p * M + m) * L + l) * Nb + t] = init_resp[p * Nb * L * M + (t * L + l) * M + m];
    CUB(cudaMemcpy(h->Q, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    const double* tr = init_resp + 4 * Nb * L * M;
    buf.assign(2 * M * Nb, 0.0);
    for (size_t X = 0; X < 2; ++X)
      for (size_t t = 0; t < Nb; ++t)
        for (size_t m = 0; m < M; ++m) buf[(X * M + m) * Nb + t] = tr[X * Nb * M + t * M + m];
    CUB(cudaMemcpy(h->QT, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  fill_kernel<<<64, 256, 0, h->st>>>(h->Wg, 2 * M * (size_t)D.F, 1.0);   // W == 1 (apvast.py:326-327)
  CUB(cudaStreamSynchronize(h->st));
#undef TRYB
#undef CUB
  *out = h;
  return OK;
}

void apv_destroy(apv_handle* h) {
  if (!h) return;
  void* ps[] = {h->rirT, h->rirTT, h->win, h->tw, h->G2, h->xin, h->Q, h->QT, h->O, h->OT, h->S, h->ST, h->Sp,
                h->Wg, h->seed, h->Pbuf, h->norms, h->pvec, h->tframe, h->tspec, h->G, h->Gt, h->R, h->rvec, h->lam, h->U, h->W, h->d_in, h->d_out,
                h->d_out_t};
  for (void* p : ps)
    if (p) cudaFree(p);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  jdiag_free(h->jd);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_syrk)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_timer)
    if (e) cudaEventDestroy(e);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
}

int apv_process_block(apv_handle* h, const double* in_A, const double* in_B, double* out_A, double* out_B,
                      double* out_A_t, double* out_B_t) {
  if (!h || !in_A || !in_B) return fail(EINVAL_, "null argument");
  if (h->cfg.perceptual == 1 && !h->G2) return fail(EINVAL_, "perceptual model tables not set (apv_set_gain_table)");
  if (h->cfg.perceptual == 2) return fail(EINVAL_, "perceptual == 2 needs apv_begin_block / apv_finish_block");
  APV_TRY(copy_in(*h, in_A, in_B));
  APV_TRY(run_block(*h, h->d_in, h->d_in + h->D.H, false, false));
  APV_TRY(copy_out(*h, out_A, out_B, out_A_t, out_B_t));
  return check_info(*h);
}

int apv_process_blocks(apv_handle* h, int nblocks, const double* in_A, const double* in_B, double* out_A,
                       double* out_B, double* out_A_t, double* out_B_t, double* w_out) {
  if (!h || nblocks < 0) return fail(EINVAL_, "bad argument");
  const Dims& D = h->D;
  const size_t per = (size_t)D.V * D.H * D.L, pert = (size_t)D.H * D.L;
  for (int b = 0; b < nblocks; ++b) {
    APV_TRY(apv_process_block(h, in_A + (size_t)b * D.H, in_B + (size_t)b * D.H, out_A ? out_A + b * per : nullptr,
                              out_B ? out_B + b * per : nullptr, out_A_t ? out_A_t + b * pert : nullptr,
                              out_B_t ? out_B_t + b * pert : nullptr));
    if (w_out) APV_TRY(apv_get(h, APV_T_W, w_out + (size_t)b * 2 * D.V * D.n, 2 * (size_t)D.V * D.n));
  }
  return OK;
}

int apv_process_block_device(apv_handle* h, const double* d_in_A, const double* d_in_B) {
  if (!h || !d_in_A || !d_in_B) return fail(EINVAL_, "null argument");
  if (h->cfg.perceptual == 2) return fail(EINVAL_, "perceptual == 2 needs apv_begin_block / apv_finish_block");
  return run_block(*h, d_in_A, d_in_B, false, false);
}

int apv_begin_block(apv_handle* h, const double* in_A, const double* in_B) {
  if (!h || !in_A || !in_B) return fail(EINVAL_, "null argument");
  APV_TRY(copy_in(*h, in_A, in_B));
  h->launches = 0;
  APV_TRY(stage_fir(*h, h->d_in, h->d_in + h->D.H));
  APV_TRY(stage_targets(*h, true));        // windowed target frames for the host gain model
  h->began = true;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_finish_block(apv_handle* h, double* out_A, double* out_B, double* out_A_t, double* out_B_t) {
  if (!h || !h->began) return fail(EINVAL_, "apv_finish_block without apv_begin_block");
  h->began = false;
  const int saved = h->launches;
  APV_TRY(run_block(*h, nullptr, nullptr, true, false));
  h->launches += saved;
  APV_TRY(copy_out(*h, out_A, out_B, out_A_t, out_B_t));
  return check_info(*h);
}

int apv_advance_state(apv_handle* h, const double* in_A, const double* in_B) {
  if (!h || !in_A || !in_B) return fail(EINVAL_, "null argument");
  if (h->cfg.perceptual == 2) return fail(EINVAL_, "perceptual == 2 needs apv_begin_block / apv_finish_block");
  APV_TRY(copy_in(*h, in_A, in_B));
  APV_TRY(run_block(*h, h->d_in, h->d_in + h->D.H, false, true));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_get(apv_handle* h, int id, double* dst, size_t count) {
  if (!h || !dst) return fail(EINVAL_, "null argument");
  TensorInfo ti = tensor_info(*h, id);
  if (!ti.ptr || count != ti.count) return fail(EINVAL_, "apv_get: bad tensor id %d or count %zu (want %zu)", id, count, ti.count);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  if (id == APV_T_R) {
    const Dims& D = h->D;
    APV_CUDA_TRY(cudaMemcpy2D(dst, (size_t)D.n * sizeof(double), ti.ptr, (size_t)D.ldn * sizeof(double),
                              (size_t)D.n * sizeof(double), (size_t)4 * D.n, cudaMemcpyDeviceToHost));
  } else {
    APV_CUDA_TRY(cudaMemcpy(dst, ti.ptr, count * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return OK;
}

int apv_set(apv_handle* h, int id, const double* src, size_t count) {
  if (!h || !src) return fail(EINVAL_, "null argument");
  TensorInfo ti = tensor_info(*h, id);
  if (!ti.ptr || count != ti.count) return fail(EINVAL_, "apv_set: bad tensor id %d or count %zu (want %zu)", id, count, ti.count);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  if (id == APV_T_R) {
    const Dims& D = h->D;
    APV_CUDA_TRY(cudaMemcpy2D(ti.ptr, (size_t)D.ldn * sizeof(double), src, (size_t)D.n * sizeof(double),
                              (size_t)D.n * sizeof(double), (size_t)4 * D.n, cudaMemcpyHostToDevice));
  } else {
    APV_CUDA_TRY(cudaMemcpy(ti.ptr, src, count * sizeof(double), cudaMemcpyHostToDevice));
  }
  return OK;
}

int apv_set_mu(apv_handle* h, double mu) {
  if (!h) return fail(EINVAL_, "null argument");
  h->cfg.mu = mu;
  return OK;
}

int apv_set_gain_table(apv_handle* h, int n_channels, const double* G2, double Cs, double Ca, double Leff) {
  if (!h || !G2 || n_channels < 1) return fail(EINVAL_, "bad argument");
  if (h->G2) cudaFree(h->G2);
  h->G2 = nullptr;
  const size_t cnt = (size_t)n_channels * h->D.F;
  APV_CUDA_TRY(cudaMalloc((void**)&h->G2, cnt * sizeof(double)));
  APV_CUDA_TRY(cudaMemcpy(h->G2, G2, cnt * sizeof(double), cudaMemcpyHostToDevice));
  h->nchan = n_channels; h->Cs = Cs; h->Ca = Ca; h->Leff = Leff;
  return OK;
}

int apv_sweep(apv_handle* h, int n_mu, const double* mu, double* w_out) {
  if (!h || !mu || !w_out || n_mu < 1) return fail(EINVAL_, "bad argument");
  const Dims& D = h->D;
  const size_t cnt = 2 * (size_t)D.V * D.n;
  double* tmp = nullptr;
  APV_CUDA_TRY(cudaMalloc((void**)&tmp, cnt * sizeof(double)));
  APV_CUDA_TRY(cudaMemsetAsync(tmp, 0, cnt * sizeof(double), h->st));
  int rc = OK;
  for (int i = 0; i < n_mu && rc == OK; ++i) {
    rc = stage_sweep(*h, mu[i], tmp);
    if (rc == OK && cudaMemcpyAsync(w_out + (size_t)i * cnt, tmp, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess)
      rc = fail(ECUDA, "sweep copy failed");
  }
  cudaStreamSynchronize(h->st);
  cudaFree(tmp);
  return rc;
}

int apv_eval_zone(apv_handle* h, int zone, int n_samples, const double* feeds, const double* signal, double* out3) {
  if (!h || !feeds || !signal || !out3) return fail(EINVAL_, "null argument");
  return eval_zone(*h, zone, n_samples, feeds, signal, out3);
}

int apv_device_ptr(apv_handle* h, int id, void** ptr) {
  if (!h || !ptr) return fail(EINVAL_, "null argument");
  TensorInfo ti = tensor_info(*h, id);
  if (!ti.ptr) return fail(EINVAL_, "bad tensor id %d", id);
  *ptr = ti.ptr;
  return OK;
}

int apv_synchronize(apv_handle* h) {
  if (!h) return fail(EINVAL_, "null argument");
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_stage_times(apv_handle* h, float* ms7) {
  if (!h || !ms7) return fail(EINVAL_, "null argument");
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  for (int i = 0; i < 6; ++i) {
    ms7[i] = 0.f;
    cudaEventElapsedTime(&ms7[i], h->ev[i], h->ev[i + 1]);
  }
  ms7[6] = 0.f;
  cudaEventElapsedTime(&ms7[6], h->ev[0], h->ev[6]);
  return OK;
}

int apv_jdiag_phase_times(apv_handle* h, float* ms6) {
  if (!h || !ms6) return fail(EINVAL_, "null argument");
  for (int i = 0; i < 6; ++i) ms6[i] = 0.f;
  if (h->nz == 0) return OK;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms6[i], h->jd.ev[i], h->jd.ev[i + 1]);
  return OK;
}

int apv_kernel_times(apv_handle* h, float* ms4) {
  if (!h || !ms4) return fail(EINVAL_, "null argument");
  for (int i = 0; i < 4; ++i) ms4[i] = 0.f;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  if (h->nz > 0) {
    cudaEventElapsedTime(&ms4[1], h->ev_syrk[0], h->ev_syrk[1]);
    if (h->jd.last_two_stage) {          // band reduction | bulge chasing of the last block
      cudaEventElapsedTime(&ms4[0], h->jd.ev[2], h->jd.ev2[0]);
      cudaEventElapsedTime(&ms4[3], h->jd.ev2[0], h->jd.ev2[1]);
      ms4[2] = -1.f;
    } else if (h->jd.last_panels) {
      float tot = 0.f;
      for (int p = 0; p < h->jd.npanel; ++p) {
        float t = 0.f;
        cudaEventElapsedTime(&t, h->jd.pev[2 * p], h->jd.pev[2 * p + 1]);
        tot += t;
      }
      ms4[0] = tot;
      ms4[2] = (float)h->jd.npanel;
    }
  }
  return OK;
}

int apv_timer_start(apv_handle* h) {
  if (!h) return fail(EINVAL_, "null argument");
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  APV_CUDA_TRY(cudaEventRecord(h->ev_timer[0], h->st));
  return OK;
}

int apv_timer_stop(apv_handle* h, float* ms) {
  if (!h || !ms) return fail(EINVAL_, "null argument");
  APV_CUDA_TRY(cudaEventRecord(h->ev_timer[1], h->st));
  APV_CUDA_TRY(cudaEventSynchronize(h->ev_timer[1]));
  APV_CUDA_TRY(cudaEventElapsedTime(ms, h->ev_timer[0], h->ev_timer[1]));
  return OK;
}

int apv_launch_count(const apv_handle* h) { return h ? h->launches : 0; }

int apv_jdiag(int n, int V, const double* A, const double* B, double reg, int eig_mode, double* lambda_out,
              double* U_out, int* pivot_out) {
  if (n < 1 || V < 1 || V > n || !A || !B || !lambda_out || !U_out) return fail(EINVAL_, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ECUDA, "no CUDA device: no CPU fallback");
  JdiagWs ws;
  int rc = jdiag_alloc(ws, n, V, 1, eig_mode);
  double *dA = nullptr, *dB = nullptr;
  const size_t bytes = (size_t)n * n * sizeof(double);
  if (rc == OK && (cudaMalloc((void**)&dA, bytes) != cudaSuccess || cudaMalloc((void**)&dB, bytes) != cudaSuccess))
    rc = fail(ECUDA, "cudaMalloc failed");
  if (rc == OK) {
    cudaMemcpy(dA, A, bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B, bytes, cudaMemcpyHostToDevice);
    const double* br[2] = {dA, dA};
    const double* dk[2] = {dB, dB};
    int nl = 0;
    rc = jdiag_run(ws, br, dk, n, reg, 0, &nl);
  }
  if (rc == OK) {
    int info[4] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(ECUDA, "jdiag kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == OK) {
      cudaMemcpy(info, ws.info, sizeof(info), cudaMemcpyDeviceToHost);
      cudaMemcpy(lambda_out, ws.lam, (size_t)V * sizeof(double), cudaMemcpyDeviceToHost);
      cudaMemcpy(U_out, ws.Zt, (size_t)V * n * sizeof(double), cudaMemcpyDeviceToHost);
      if (pivot_out) *pivot_out = info[0];
      if (info[0] != 0) rc = fail(ENOTPD, "Matrix is not positive definite (pivot %d)", info[0]);
      else if (info[1] != 0) rc = fail(ENOCONV, "inverse iteration did not converge");
    }
  }
  if (dA) cudaFree(dA);
  if (dB) cudaFree(dB);
  jdiag_free(ws);
  return rc;
}

int apv_util_gemm(int M, int N, int K, int transA, int transB, double alpha, const double* A, const double* B,
                  double beta, double* C) {
  if (M < 1 || N < 1 || K < 0 || !A || !B || !C) return fail(EINVAL_, "bad argument");
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  const size_t sa = (size_t)M * K, sb = (size_t)K * N, sc = (size_t)M * N;
  APV_CUDA_TRY(cudaMalloc((void**)&dA, (sa ? sa : 1) * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dB, (sb ? sb : 1) * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dC, sc * sizeof(double)));
  cudaMemcpy(dA, A, sa * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B, sb * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C, sc * sizeof(double), cudaMemcpyHostToDevice);
  GemmArgs g{};
  g.A = dA; g.B = dB; g.C = dC; g.M = M; g.N = N; g.K = K;
  g.lda = transA ? M : K; g.ldb = transB ? K : N; g.ldc = N;
  g.alpha = alpha; g.beta = beta; g.transA = transA; g.transB = transB; g.batch = 1;
  int rc = gemm_f64(g, 0);
  if (rc == OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(ECUDA, "gemm failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == OK) cudaMemcpy(C, dC, sc * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return rc;
}

int apv_util_fft(int n, int inverse, const double* in_ri, double* out_ri) {
  if (!in_ri || !out_ri) return fail(EINVAL_, "null argument");
  return fft_util(n, inverse, in_ri, out_ri);
}

int apv_bench_gemm(int n, int nrep, float* ms) {
  if (n < 1 || nrep < 1 || !ms) return fail(EINVAL_, "bad argument");
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  const size_t cnt = (size_t)n * n;
  APV_CUDA_TRY(cudaMalloc((void**)&dA, cnt * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dB, cnt * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dC, cnt * sizeof(double)));
  fill_kernel<<<256, 256>>>(dA, cnt, 1.0 / n);
  fill_kernel<<<256, 256>>>(dB, cnt, 0.5);
  GemmArgs g{};
  g.A = dA; g.B = dB; g.C = dC; g.M = g.N = g.K = n; g.lda = g.ldb = g.ldc = n; g.alpha = 1.0; g.beta = 0.0; g.batch = 1;
  g.transB = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = gemm_f64(g, 0);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < nrep && rc == OK; ++i) rc = gemm_f64(g, 0);
  cudaEventRecord(e1, 0);
  if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(ECUDA, "gemm bench failed");
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  *ms = t / nrep;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return rc;
}

}  // extern "C"
