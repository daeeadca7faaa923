// C-ABI of the AP-VAST B200 engine (include/apvast_b200.h): handle life-cycle, the per-block call that
// chains S1..S7 on one CUDA stream, state get/set, and the test / measurement utilities.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "engine.cuh"

using namespace apv;

namespace apv {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(apv::g_err, sizeof(apv::g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace apv

namespace {

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
  return jdiag_run(h.jd, bright, dark, D.ldn, h.cfg.loading_mode == 1 ? 0.0 : h.cfg.reg, h.st, &h.launches,
                   (h.cfg.reg_relative && h.cfg.loading_mode != 1) ? h.regv : nullptr);
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

void use_slot(Handle& h, int slot) {
  h.R = h.Rslot[slot];
  h.rvec = h.rvslot[slot];
  h.xw = h.xwslot[slot];
}

}  // namespace

namespace apv {

// Front half of a block: S1-S3 (state) and S4 (statistics into the current slot), on the stream h.st points at.
// `skip_s1`: S1 was already run (split call).  `state_only`: warm-up of a block range, no statistics.
int run_front(Handle& h, const double* d_inA, const double* d_inB, bool skip_s1, bool state_only) {
  cudaEvent_t* ev = h.ev;
  NvtxRange nv_front("apv front half: S1-S4");
  APV_CUDA_TRY(cudaEventRecord(ev[0], h.st));
  { NvtxRange nv("S1 rir_conv"); if (!skip_s1) APV_TRY(stage_fir(h, d_inA, d_inB)); }
  APV_CUDA_TRY(cudaEventRecord(ev[1], h.st));
  {
    NvtxRange nv("S2+S3 wola_weight");
    if (!(skip_s1 && h.cfg.perceptual == 3)) APV_TRY(stage_targets(h, false));   // 3: S2 ran in apv_begin_block
    APV_TRY(stage_weighted(h));
  }
  APV_CUDA_TRY(cudaEventRecord(ev[2], h.st));
  if (!state_only) {
    NvtxRange nv("S4 stats_syrk");
    APV_TRY(stage_stats(h));
    if (h.cfg.loading_mode == 1) APV_TRY(stage_loading(h));
    else if (h.cfg.reg_relative && h.nz > 0) APV_TRY(stage_spectral_norms(h));
  }
  APV_CUDA_TRY(cudaEventRecord(ev[3], h.st));
  return OK;
}

// Back half: S5 (joint diagonalisation of the current slot), S6 (filter sum into h.W), S7 (rendering into h.d_out).
int run_back(Handle& h, cudaEvent_t order) {
  cudaEvent_t* ev = h.ev;
  NvtxRange nv_back("apv back half: S5-S7");
  APV_CUDA_TRY(cudaEventRecord(ev[7], h.st));
  { NvtxRange nv("S5 jdiag"); APV_TRY(run_jdiag(h)); }
  // with two back halves in flight the joint diagonalisations overlap, but S6 / S7 touch sequential state (the
  // published eigenpairs, the output overlap buffers G): they follow the previous block's
  if (order) APV_CUDA_TRY(cudaStreamWaitEvent(h.st, order, 0));
  APV_TRY(publish_eig(h));
  APV_CUDA_TRY(cudaEventRecord(ev[4], h.st));
  { NvtxRange nv("S6 vast_sweep"); APV_TRY(stage_sweep(h, h.cfg.mu, h.W)); }
  APV_CUDA_TRY(cudaEventRecord(ev[5], h.st));
  { NvtxRange nv("S7 render_ola"); APV_TRY(stage_render(h)); }
  APV_CUDA_TRY(cudaEventRecord(ev[6], h.st));
  return OK;
}

}  // namespace apv

namespace {

// S1..S7 in order on the main stream with the inputs already on the device (the per-block call).
int run_block(Handle& h, const double* d_inA, const double* d_inB, bool skip_s1, bool state_only) {
  h.launches = 0;
  APV_TRY(run_front(h, d_inA, d_inB, skip_s1, state_only));
  if (!state_only) {
    APV_TRY(run_back(h));
  } else {
    for (int i : {7, 4, 5, 6}) APV_CUDA_TRY(cudaEventRecord(h.ev[i], h.st));
  }
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

namespace apv {

bool pipelined(const Handle& h) {
  // the MATLAB loading and the norm-relative regularisation read spectral norms back to the host inside S4
  return h.pipeline != 0 && h.cfg.loading_mode != 1 && !h.cfg.reg_relative && h.cfg.perceptual < 2;
}

// Multi-block calls run every block in two halves.  With pipelining the front half (S1-S4) goes to the low-priority
// stream and fills slot b & 1 of the statistics while the back half (S5-S7) of block b - 1 still runs on the main
// stream: S4 of block b + 1 depends on the streaming state only (apvast.py:329-364), never on the filters of block b.
// Measured at cfg-3 (profiles/r02_pipeline_timeline_v1.txt): when the front half starts together with the back half,
// the ~900 short launches of S5 keep waiting for SMs held by SYRK CTAs and the overlap gains only 5 %.  With
// pipeline == 2 (default) the front half of block b + 1 is therefore gated on the START OF THE BULGE CHASING of block
// b (event ev2[0] of the two-stage tridiagonalisation): from there on S5 is a handful of latency-bound kernels that
// leave most SMs idle, and the statistics fill them.  The caller enqueues back(b) BEFORE front(b + 1) so that the
// event is recorded when the wait is enqueued.  `b` counts from 0 inside the call.
int enqueue_front(Handle& h, long b, const double* d_inA, const double* d_inB, bool state_only) {
  const int slot = (int)(b % h.nslot);
  const bool pipe = pipelined(h) && !state_only;
  const bool two = pipe && h.depth >= 2;
  cudaStream_t main_st = h.st;
  h.launches_front = 0;
  const int saved = h.launches;
  h.launches = 0;
  use_slot(h, slot);
  if (pipe) {
    if (b >= h.nslot) {
      APV_CUDA_TRY(cudaStreamWaitEvent(h.st_front, h.ev_free[slot], 0));     // back half of block b - nslot released the slot
    } else if (b == 0) {
      APV_CUDA_TRY(cudaEventRecord(h.ev_join, main_st));                     // state written by earlier calls
      APV_CUDA_TRY(cudaStreamWaitEvent(h.st_front, h.ev_join, 0));
      for (int k = 0; two && k < h.depth - 1; ++k) APV_CUDA_TRY(cudaStreamWaitEvent(h.st_backx[k], h.ev_join, 0));
    }
    if (b >= 1 && h.pipeline == 2 && !two && h.nz > 0 && h.jd.last_two_stage)
      APV_CUDA_TRY(cudaStreamWaitEvent(h.st_front, h.jd.ev2[0], 0));         // block b - 1 has reached its bulge chasing
    if (b >= 2 && !two)
      APV_CUDA_TRY(cudaStreamWaitEvent(h.st_front, h.ev_free[(b - 2) % h.nslot], 0));   // depth 1: at most one block ahead
    h.st = h.st_front;
  }
  if (h.dbg_ev && !state_only && b < h.dbg_cap) cudaEventRecord(h.dbg_ev[b * 4 + 0], h.st);
  int rc = run_front(h, d_inA, d_inB, false, state_only);
  if (h.dbg_ev && !state_only && b < h.dbg_cap) cudaEventRecord(h.dbg_ev[b * 4 + 1], h.st);
  h.st = main_st;
  h.launches_front = h.launches;
  h.launches = saved;
  APV_TRY(rc);
  if (pipe) APV_CUDA_TRY(cudaEventRecord(h.ev_ready[slot], h.st_front));
  return OK;
}

// Back half of block b; results go to the device buffers of `sink`; `done` (optional) is recorded behind it.
// depth 2 (default for n < 2048, where S5 is latency-bound and leaves most of the chip idle): the back halves of
// consecutive blocks alternate between two streams and two joint-diagonalisation workspaces, so two S5 run side by
// side; S6 / S7 stay in block order (ev_order).
int enqueue_back(Handle& h, long b, const BlockSink& sink, cudaEvent_t done) {
  const int slot = (int)(b % h.nslot);
  const bool pipe = pipelined(h);
  const bool two = pipe && h.depth >= 2;
  const int which = two ? (int)(b % h.depth) : 0;
  cudaStream_t main_st = h.st;
  cudaStream_t sb = which ? h.st_backx[which - 1] : main_st;
  use_slot(h, slot);
  h.launches = h.launches_front;
  if (pipe) APV_CUDA_TRY(cudaStreamWaitEvent(sb, h.ev_ready[slot], 0));
  h.W = sink.W ? sink.W : h.home_W;
  h.d_out = sink.out ? sink.out : h.home_out;
  h.d_out_t = sink.out_t ? sink.out_t : h.home_out_t;
  if (which) std::swap(h.jd, h.jdx[which - 1]);
  h.st = sb;
  if (h.dbg_ev && b < h.dbg_cap) cudaEventRecord(h.dbg_ev[b * 4 + 2], sb);
  int rc = run_back(h, (two && b >= 1) ? h.ev_order : nullptr);
  if (h.dbg_ev && b < h.dbg_cap) cudaEventRecord(h.dbg_ev[b * 4 + 3], sb);
  if (rc == OK && two) rc = cudaEventRecord(h.ev_order, sb) == cudaSuccess ? OK : fail(ECUDA, "event record failed");
  if (rc == OK && sink.info && h.nz > 0)
    rc = cudaMemcpyAsync(sink.info, h.jd.info, (size_t)h.nz * 4 * sizeof(int), cudaMemcpyDeviceToDevice, sb) == cudaSuccess
             ? OK : fail(ECUDA, "status copy failed");
  if (rc == OK && pipe) rc = cudaEventRecord(h.ev_free[slot], sb) == cudaSuccess ? OK : fail(ECUDA, "event record failed");
  if (rc == OK && done) rc = cudaEventRecord(done, sb) == cudaSuccess ? OK : fail(ECUDA, "event record failed");
  h.st = main_st;
  if (which) std::swap(h.jd, h.jdx[which - 1]);
  h.last_ws = which;
  return rc;
}

// D back halves in flight need D joint-diagonalisation workspaces and D + 1 statistics slots; the extra ones are
// allocated by the first multi-block call (depth < 0: keep the wanted depth, just make sure they exist), so per-block
// users never pay for them.
int ensure_depth(Handle& h, int depth) {
  if (depth >= 0) {
    h.depth = h.nz > 0 ? std::max(1, std::min(depth, (int)Handle::MAXDEPTH)) : 1;
    return OK;
  }
  const Dims& D = h.D;
  for (int k = 0; k < h.depth - 1; ++k)
    if (h.jdx[k].n == 0) APV_TRY(jdiag_alloc(h.jdx[k], D.n, D.V, h.nz, h.cfg.eig_mode));
  const int want = std::max(2, h.depth + 1);
  for (int s = 0; s < want; ++s) {
    if (h.Rslot[s]) continue;
    APV_CUDA_TRY(cudaMalloc((void**)&h.Rslot[s], 4 * (size_t)D.n * D.ldn * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.rvslot[s], 2 * (size_t)D.n * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.xwslot[s], 2 * (size_t)D.Nb * sizeof(double)));
    APV_CUDA_TRY(cudaMemsetAsync(h.Rslot[s], 0, 4 * (size_t)D.n * D.ldn * sizeof(double), h.st));
    APV_CUDA_TRY(cudaMemsetAsync(h.rvslot[s], 0, 2 * (size_t)D.n * sizeof(double), h.st));
    APV_CUDA_TRY(cudaMemsetAsync(h.xwslot[s], 0, 2 * (size_t)D.Nb * sizeof(double), h.st));
  }
  h.nslot = want;
  return OK;
}

// after a multi-block call: the filters / outputs of the last block back into the handle's own buffers, so that
// apv_get and the next per-block call see the usual layout
int leave_multiblock(Handle& h) {
  const Dims& D = h.D;
  if (h.depth >= 2 && pipelined(h)) {          // the other back-half streams join the main stream
    for (int k = 0; k < h.depth - 1; ++k) {
      APV_CUDA_TRY(cudaEventRecord(h.ev_join, h.st_backx[k]));
      APV_CUDA_TRY(cudaStreamWaitEvent(h.st, h.ev_join, 0));
    }
    if (h.last_ws > 0) std::swap(h.jd, h.jdx[h.last_ws - 1]);   // h.jd = the workspace of the most recent block (apv_sweep, timers)
    h.last_ws = 0;
  }
  if (h.W != h.home_W)
    APV_CUDA_TRY(cudaMemcpyAsync(h.home_W, h.W, 2 * (size_t)D.V * D.n * sizeof(double), cudaMemcpyDeviceToDevice, h.st));
  if (h.d_out != h.home_out)
    APV_CUDA_TRY(cudaMemcpyAsync(h.home_out, h.d_out, 2 * (size_t)D.V * D.H * D.L * sizeof(double), cudaMemcpyDeviceToDevice, h.st));
  if (h.d_out_t != h.home_out_t)
    APV_CUDA_TRY(cudaMemcpyAsync(h.home_out_t, h.d_out_t, 2 * (size_t)D.H * sizeof(double), cudaMemcpyDeviceToDevice, h.st));
  h.W = h.home_W; h.d_out = h.home_out; h.d_out_t = h.home_out_t;
  return OK;
}

int status_from_info(const Handle& h, const int* info, long block) {
  for (int zi = 0; zi < h.nz; ++zi) {
    if (info[zi * 4] != 0)
      return fail(ENOTPD, "Matrix is not positive definite (block %ld, zone %c, pivot %d)", block,
                  h.zones[zi] == 0 ? 'A' : 'B', info[zi * 4]);
    if (info[zi * 4 + 1] != 0)
      return fail(ENOCONV, "eigen-solver did not converge (block %ld, zone %c, flags %d)", block,
                  h.zones[zi] == 0 ? 'A' : 'B', info[zi * 4 + 1]);
  }
  return OK;
}

}  // namespace apv

namespace {

size_t ring_slot_doubles(const Dims& D) {
  return 2 * (size_t)D.V * D.H * D.L + 2 * (size_t)D.H + 2 * (size_t)D.V * D.n + 8;   // out | out_t | W | status
}

int ensure_ring(Handle& h, int cap) {
  if (h.ring_cap >= cap) return OK;
  if (h.ring) cudaFree(h.ring);
  if (h.ring_pin) cudaFreeHost(h.ring_pin);
  h.ring = h.ring_pin = nullptr;
  h.ring_cap = h.ring_pin_cap = 0;
  const size_t sd = ring_slot_doubles(h.D);
  APV_CUDA_TRY(cudaMalloc((void**)&h.ring, (size_t)cap * sd * sizeof(double)));
  APV_CUDA_TRY(cudaMemsetAsync(h.ring, 0, (size_t)cap * sd * sizeof(double), h.st));
  APV_CUDA_TRY(cudaMallocHost((void**)&h.ring_pin, (size_t)cap * sd * sizeof(double)));
  h.ring_cap = h.ring_pin_cap = cap;
  return OK;
}

}  // namespace

extern "C" {

const char* apv_last_error(void) { return apv::g_err; }
const char* apv_version(void) { return "apvast_b200 0.3 (sm_100a)"; }

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
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return bail(fail(ECUDA, "no CUDA device: the AP-VAST B200 engine has no CPU fallback"));
  int cur = 0;
  cudaGetDevice(&cur);
  h->device = cfg->device >= 0 ? cfg->device : cur;
  if (h->device >= ndev) return bail(fail(ECUDA, "cudaSetDevice(%d) failed: %d device(s) visible", h->device, ndev));
  DevGuard dg(h->device);      // everything below, and every later call on the handle, runs on the handle's device
#define TRYB(x) do { int _s = (x); if (_s != OK) return bail(_s); } while (0)
#define CUB(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return bail(fail(ECUDA, "%s -> %s", #x, cudaGetErrorString(_e))); } while (0)
  {
    int lo = 0, hi = 0;      // numerically lowest value = highest priority
    CUB(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUB(cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, hi));
    CUB(cudaStreamCreateWithPriority(&h->st_front, cudaStreamNonBlocking, lo));
    CUB(cudaStreamCreateWithPriority(&h->st_copy, cudaStreamNonBlocking, hi));
    for (auto& sx : h->st_backx) CUB(cudaStreamCreateWithPriority(&sx, cudaStreamNonBlocking, hi));
  }
  for (auto& e : h->ev) CUB(cudaEventCreate(&e));
  for (auto& e : h->ev_ready) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : h->ev_free) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUB(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
  CUB(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  for (auto& e : h->ev_rend) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : h->ev_d2h) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  {
    const char* pe = getenv("APV_PIPELINE");
    h->pipeline = pe ? atoi(pe) : 2;
  }
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
  if (cfg->stats_mode != 2) {
    // microphones per SYRK launch (their partial tiles of one (tile, path) share a ring slot) and ring length
    int G = std::min(16, (int)M);
    if (const char* e = getenv("APV_SYRK_GROUP")) G = std::max(1, std::min(16, atoi(e)));
    h->syrk_group = G;
    if (const char* e = getenv("APV_SYRK_SLOTS")) h->syrk_slots = std::max(2, atoi(e));
    TRYB(dalloc(&h->Pbuf, (size_t)h->syrk_slots * 16 * 128 * 128));
    const size_t nt = (n + 127) / 128;
    TRYB(dalloc(&h->syrk_cnt, 4 * (nt * (nt + 1) / 2) + (size_t)h->syrk_slots));
  }
  TRYB(dalloc(&h->norms, 64));
  TRYB(dalloc(&h->pvec, 8 * n));
  TRYB(dalloc(&h->tframe, 2 * M * Nb));
  TRYB(dalloc(&h->G, 2 * V * L * Nb));
  TRYB(dalloc(&h->Gt, 2 * Nb));
  for (int s = 0; s < 2; ++s) {              // (further slots: ensure_depth, on the first multi-block call)
    TRYB(dalloc(&h->Rslot[s], 4 * n * (size_t)D.ldn));
    TRYB(dalloc(&h->rvslot[s], 2 * n));
    TRYB(dalloc(&h->xwslot[s], 2 * Nb));
  }
  h->R = h->Rslot[0]; h->rvec = h->rvslot[0]; h->xw = h->xwslot[0];
  TRYB(dalloc(&h->regv, 2));
  TRYB(dalloc(&h->lam, 2 * V));
  TRYB(dalloc(&h->U, 2 * V * n));
  TRYB(dalloc(&h->home_W, 2 * V * n));
  TRYB(dalloc(&h->d_in, 2 * (size_t)D.H));
  TRYB(dalloc(&h->home_out, 2 * V * (size_t)D.H * L));
  TRYB(dalloc(&h->home_out_t, 2 * (size_t)D.H));
  h->W = h->home_W; h->d_out = h->home_out; h->d_out_t = h->home_out_t;
  h->nz = 0;
  if (D.runA) h->zones[h->nz++] = 0;
  if (D.runB) h->zones[h->nz++] = 1;
  if (h->nz == 1) h->zones[1] = h->zones[0];
  if (h->nz > 0) TRYB(jdiag_alloc(h->jd, D.n, D.V, h->nz, cfg->eig_mode));
  {
    // joint diagonalisations in flight in multi-block calls: 3 from n = 2048 (measured at cfg-3: one 110.5, two 97.4,
    // three 94.5, four 94.4 ms per block -- with the latency-bound kernels of S5 on fewer SMs since the second session
    // of round 2 a third one pays; in the first session it did not: 99.2 vs 97.6), 4 below n = 2048 where S5 is
    // latency-bound throughout; APV_DEPTH overrides
    const char* de = getenv("APV_DEPTH");
    TRYB(ensure_depth(*h, de ? atoi(de) : (D.n < 2048 ? 4 : 3)));
  }
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
    // silent padding microphones of zone A (active_mics_A, multi-zone composition) do not exist: their start buffers
    // stay zero, whatever the statistics kernels sum over
    const size_t MA = (cfg->active_mics_A > 0 && cfg->active_mics_A < D.M) ? (size_t)cfg->active_mics_A : M;
    buf.assign(4 * M * L * Nb, 0.0);
    for (size_t p = 0; p < 4; ++p)
      for (size_t t = 0; t < Nb; ++t)
        for (size_t l = 0; l < L; ++l)
          for (size_t m = 0; m < ((p & 1) == 0 ? MA : M); ++m)       // paths 0 (A->A) and 2 (B->A) end in zone A
            buf[((p * M + m) * L + l) * Nb + t] = init_resp[p * Nb * L * M + (t * L + l) * M + m];
    CUB(cudaMemcpy(h->Q, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    const double* tr = init_resp + 4 * Nb * L * M;
    buf.assign(2 * M * Nb, 0.0);
    for (size_t X = 0; X < 2; ++X)
      for (size_t t = 0; t < Nb; ++t)
        for (size_t m = 0; m < (X == 0 ? MA : M); ++m) buf[(X * M + m) * Nb + t] = tr[X * Nb * M + t * M + m];
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
  DevGuard dg(h->device);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->st_front) cudaStreamSynchronize(h->st_front);
  if (h->st_copy) cudaStreamSynchronize(h->st_copy);
  range_free(*h);
  void* ps[] = {h->rirT, h->rirTT, h->win, h->tw, h->G2, h->xin, h->Q, h->QT, h->O, h->OT, h->S, h->ST, h->Sp,
                h->Wg, h->seed, h->Pbuf, h->syrk_cnt, h->norms, h->pvec, h->tframe, h->tspec, h->G, h->Gt, h->regv,
                h->lam, h->U, h->home_W, h->d_in, h->home_out, h->home_out_t, h->ring};
  for (void* p : ps)
    if (p) cudaFree(p);
  for (int s = 0; s < Handle::NSLOT; ++s)
    for (void* p : {(void*)h->Rslot[s], (void*)h->rvslot[s], (void*)h->xwslot[s]})
      if (p) cudaFree(p);
  if (h->sw_mu) cudaFree(h->sw_mu);
  if (h->sw_m) cudaFree(h->sw_m);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->ring_pin) cudaFreeHost(h->ring_pin);
  if (h->ring_info) cudaFreeHost(h->ring_info);
  jdiag_free(h->jd);
  for (auto& w : h->jdx) jdiag_free(w);
  for (cudaEvent_t* arr : {h->ev_ready, h->ev_free})
    for (int i = 0; i < Handle::NSLOT; ++i)
      if (arr[i]) cudaEventDestroy(arr[i]);
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (auto& sx : h->st_backx)
    if (sx) { cudaStreamSynchronize(sx); cudaStreamDestroy(sx); }
  for (cudaEvent_t* arr : {h->ev_rend, h->ev_d2h})
    for (int i = 0; i < 8; ++i)
      if (arr[i]) cudaEventDestroy(arr[i]);
  if (h->st_front) cudaStreamDestroy(h->st_front);
  if (h->st_copy) cudaStreamDestroy(h->st_copy);
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
  if (h->cfg.perceptual >= 2) return fail(EINVAL_, "this perceptual mode needs apv_begin_block / apv_finish_block");
  DevGuard dg(h->device);
  APV_TRY(copy_in(*h, in_A, in_B));
  APV_TRY(run_block(*h, h->d_in, h->d_in + h->D.H, false, false));
  APV_TRY(copy_out(*h, out_A, out_B, out_A_t, out_B_t));
  return check_info(*h);
}

// Throughput path.  The hops are copied to the device once; block b + 1's S1-S4 overlap block b's S5-S7 on a second
// stream (two statistics slots); every block is rendered into a ring slot in HBM, copied to a pinned host ring on a
// copy stream and handed to the caller one block behind, so nothing on the device waits for the host.
int apv_process_blocks(apv_handle* h, int nblocks, const double* in_A, const double* in_B, double* out_A,
                       double* out_B, double* out_A_t, double* out_B_t, double* w_out) {
  if (!h || nblocks < 0 || (nblocks > 0 && (!in_A || !in_B))) return fail(EINVAL_, "bad argument");
  if (h->cfg.perceptual == 1 && !h->G2) return fail(EINVAL_, "perceptual model tables not set (apv_set_gain_table)");
  if (h->cfg.perceptual >= 2) return fail(EINVAL_, "this perceptual mode needs apv_begin_block / apv_finish_block");
  if (nblocks == 0) return OK;
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const size_t per = (size_t)D.V * D.H * D.L, pert = (size_t)D.H * D.L, perw = 2 * (size_t)D.V * D.n;
  const size_t sd = ring_slot_doubles(D);
  APV_TRY(ensure_depth(*h, -1));
  // the host hands blocks to the caller `lag` blocks behind its own enqueueing, so that `depth` back halves can be in
  // flight on the device; the rings hold the blocks in between
  const int lag = sd * sizeof(double) > ((size_t)256 << 20) ? 1 : std::max(1, std::min(h->depth, 6));
  const int cap = lag + 1 + (lag > 1 ? 1 : (sd * sizeof(double) > ((size_t)256 << 20) ? 0 : 1));
  APV_TRY(ensure_ring(*h, cap));
  double* d_sig = nullptr;       // [2][nblocks][H]
  const size_t sig = (size_t)nblocks * D.H;
  APV_CUDA_TRY(cudaMalloc((void**)&d_sig, 2 * sig * sizeof(double)));
  int rc = OK;
  auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == OK) rc = fail(ECUDA, "%s -> %s", what, cudaGetErrorString(e)); };
  cu(cudaMemcpyAsync(d_sig, in_A, sig * sizeof(double), cudaMemcpyHostToDevice, h->st), "H2D in_A");
  cu(cudaMemcpyAsync(d_sig + sig, in_B, sig * sizeof(double), cudaMemcpyHostToDevice, h->st), "H2D in_B");
  const int ref = D.refA, lt = (D.J * ref + D.d) / D.J;
  const int ltB = h->cfg.target_ref_per_zone ? (D.J * D.refB + D.d) / D.J : lt;
  auto retire = [&](long b) {
    const int rs = (int)(b % cap);
    cu(cudaEventSynchronize(h->ev_d2h[rs]), "wait D2H");
    if (rc != OK) return;
    const double* src = h->ring_pin + (size_t)rs * sd;
    const double* st_ = src + 2 * per + 2 * (size_t)D.H + perw;
    int info[8];
    memcpy(info, st_, sizeof(info));
    const int s2 = status_from_info(*h, info, b);
    if (s2 != OK) { rc = s2; return; }
    if (out_A && D.runA) memcpy(out_A + (size_t)b * per, src, per * sizeof(double));
    if (out_B && D.runB) memcpy(out_B + (size_t)b * per, src + per, per * sizeof(double));
    const double* t = src + 2 * per;
    for (int X = 0; X < 2; ++X) {
      double* o = X == 0 ? out_A_t : out_B_t;
      if (!o) continue;
      o += (size_t)b * pert;
      memset(o, 0, pert * sizeof(double));
      const int col = X == 0 ? lt : ltB;
      for (int i = 0; i < D.H; ++i) o[(size_t)i * D.L + col] = t[(size_t)X * D.H + i];
    }
    if (w_out) memcpy(w_out + (size_t)b * perw, src + 2 * per + 2 * (size_t)D.H, perw * sizeof(double));
  };
  auto sig_ptr = [&](int X, long b) { return d_sig + (size_t)X * sig + (size_t)b * D.H; };
  if (rc == OK) rc = enqueue_front(*h, 0, sig_ptr(0, 0), sig_ptr(1, 0), false);
  for (long b = 0; b < nblocks && rc == OK; ++b) {
    const int rs = (int)(b % cap);
    double* slot = h->ring + (size_t)rs * sd;
    if (b >= cap) {               // the slot's previous contents have left for the host (either back-half stream may render next)
      cu(cudaStreamWaitEvent(h->st, h->ev_d2h[rs], 0), "wait ring slot");
      for (auto& sx : h->st_backx) cu(cudaStreamWaitEvent(sx, h->ev_d2h[rs], 0), "wait ring slot");
    }
    BlockSink sink{slot, slot + 2 * per, slot + 2 * per + 2 * (size_t)D.H,
                   reinterpret_cast<int*>(slot + 2 * per + 2 * (size_t)D.H + perw)};
    if (rc == OK) rc = enqueue_back(*h, b, sink, h->ev_rend[rs]);
    if (rc == OK && b + 1 < nblocks) rc = enqueue_front(*h, b + 1, sig_ptr(0, b + 1), sig_ptr(1, b + 1), false);
    cu(cudaStreamWaitEvent(h->st_copy, h->ev_rend[rs], 0), "wait render");
    cu(cudaMemcpyAsync(h->ring_pin + (size_t)rs * sd, slot, sd * sizeof(double), cudaMemcpyDeviceToHost, h->st_copy), "D2H");
    cu(cudaEventRecord(h->ev_d2h[rs], h->st_copy), "record");
    if (b >= lag && rc == OK) retire(b - lag);
  }
  for (long b = std::max(0L, (long)nblocks - lag); b < nblocks && rc == OK; ++b) retire(b);
  const int rc2 = leave_multiblock(*h);
  cudaStreamSynchronize(h->st_front);
  cudaStreamSynchronize(h->st_copy);
  for (auto& sx : h->st_backx) cudaStreamSynchronize(sx);
  cudaStreamSynchronize(h->st);
  cudaFree(d_sig);
  return rc != OK ? rc : rc2;
}

int apv_process_block_device(apv_handle* h, const double* d_in_A, const double* d_in_B) {
  if (!h || !d_in_A || !d_in_B) return fail(EINVAL_, "null argument");
  if (h->cfg.perceptual >= 2) return fail(EINVAL_, "this perceptual mode needs apv_begin_block / apv_finish_block");
  DevGuard dg(h->device);
  return run_block(*h, d_in_A, d_in_B, false, false);
}

// perceptual == 2: S1 + the windowed target frames for a host gain model; perceptual == 3: S1 + S2 with the on-device
// masking model, after which the caller may overwrite weighting curves of other microphone groups (apv_copy_weights).
int apv_begin_block(apv_handle* h, const double* in_A, const double* in_B) {
  if (!h || !in_A || !in_B) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  APV_TRY(copy_in(*h, in_A, in_B));
  h->launches = 0;
  APV_TRY(stage_fir(*h, h->d_in, h->d_in + h->D.H));
  APV_TRY(stage_targets(*h, h->cfg.perceptual != 3));   // 2: frames only; 3: the whole of S2
  h->began = true;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_finish_block(apv_handle* h, double* out_A, double* out_B, double* out_A_t, double* out_B_t) {
  if (!h || !h->began) return fail(EINVAL_, "apv_finish_block without apv_begin_block");
  DevGuard dg(h->device);
  h->began = false;
  const int saved = h->launches;
  APV_TRY(run_block(*h, nullptr, nullptr, true, false));
  h->launches += saved;
  APV_TRY(copy_out(*h, out_A, out_B, out_A_t, out_B_t));
  return check_info(*h);
}

int apv_copy_weights(apv_handle* dst, int dst_zone, int dst_mic0, apv_handle* src, int src_zone, int src_mic0, int n_mics) {
  if (!dst || !src || n_mics < 0) return fail(EINVAL_, "bad argument");
  if (dst->D.F != src->D.F) return fail(EINVAL_, "apv_copy_weights: block sizes differ");
  if (dst_zone < 0 || dst_zone > 1 || src_zone < 0 || src_zone > 1 || dst_mic0 < 0 || src_mic0 < 0 ||
      dst_mic0 + n_mics > dst->D.M || src_mic0 + n_mics > src->D.M)
    return fail(EINVAL_, "apv_copy_weights: microphone range out of bounds");
  if (dst->device != src->device) return fail(EINVAL_, "apv_copy_weights: handles on different devices");
  DevGuard dg(dst->device);
  const size_t F = dst->D.F;
  APV_CUDA_TRY(cudaStreamSynchronize(src->st));
  APV_CUDA_TRY(cudaMemcpyAsync(dst->Wg + ((size_t)dst_zone * dst->D.M + dst_mic0) * F,
                               src->Wg + ((size_t)src_zone * src->D.M + src_mic0) * F, (size_t)n_mics * F * sizeof(double),
                               cudaMemcpyDeviceToDevice, dst->st));
  return OK;
}

int apv_advance_state(apv_handle* h, const double* in_A, const double* in_B) {
  if (!h || !in_A || !in_B) return fail(EINVAL_, "null argument");
  if (h->cfg.perceptual >= 2) return fail(EINVAL_, "this perceptual mode needs apv_begin_block / apv_finish_block");
  DevGuard dg(h->device);
  APV_TRY(copy_in(*h, in_A, in_B));
  APV_TRY(run_block(*h, h->d_in, h->d_in + h->D.H, false, true));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

/* Diagnostic: timeline of the next multi-block call.  apv_debug_timeline(h, nblocks, NULL) arms the recording of four
 * events per block (front start / front end / back start / back end); a later call with `ms` returns their times in
 * milliseconds relative to the first one ((nblocks, 4), -1 where nothing was recorded). */
int apv_debug_timeline(apv_handle* h, int nblocks, float* ms) {
  if (!h || nblocks < 1) return fail(EINVAL_, "bad argument");
  DevGuard dg(h->device);
  if (!ms) {
    if (h->dbg_ev) {
      for (int i = 0; i < 4 * h->dbg_cap; ++i) cudaEventDestroy(h->dbg_ev[i]);
      delete[] h->dbg_ev;
    }
    h->dbg_ev = new cudaEvent_t[4 * (size_t)nblocks];
    h->dbg_cap = nblocks;
    for (int i = 0; i < 4 * nblocks; ++i) APV_CUDA_TRY(cudaEventCreate(&h->dbg_ev[i]));
    return OK;
  }
  if (!h->dbg_ev) return fail(EINVAL_, "apv_debug_timeline: not armed");
  APV_CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < 4 * std::min(nblocks, h->dbg_cap); ++i) {
    float t = -1.f;
    if (cudaEventElapsedTime(&t, h->dbg_ev[0], h->dbg_ev[i]) != cudaSuccess) { t = -1.f; cudaGetLastError(); }
    ms[i] = t;
  }
  return OK;
}

int apv_set_pipeline(apv_handle* h, int on) {
  if (!h) return fail(EINVAL_, "null argument");
  h->pipeline = on < 0 ? 0 : (on > 2 ? 2 : on);
  return OK;
}

int apv_set_depth(apv_handle* h, int depth) {
  if (!h) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return ensure_depth(*h, depth);
}

int apv_set_reg_mode(apv_handle* h, int relative) {
  if (!h) return fail(EINVAL_, "null argument");
  h->cfg.reg_relative = relative != 0;
  return OK;
}

int apv_get(apv_handle* h, int id, double* dst, size_t count) {
  if (!h || !dst) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
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
  DevGuard dg(h->device);
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
  DevGuard dg(h->device);
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
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const size_t cnt = 2 * (size_t)D.V * D.n;
  // chunks of mu values so that the device buffer stays below ~1 GB
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_mu, ((size_t)1 << 30) / (cnt * sizeof(double))));
  double *tmp = nullptr, *d_mu = nullptr;
  APV_CUDA_TRY(cudaMalloc((void**)&tmp, (size_t)chunk * cnt * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&d_mu, n_mu * sizeof(double)));
  int rc = OK;
  if (cudaMemcpyAsync(d_mu, mu, n_mu * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess) rc = fail(ECUDA, "mu copy failed");
  for (int i = 0; i < n_mu && rc == OK; i += chunk) {
    const int c = std::min(chunk, n_mu - i);
    cudaMemsetAsync(tmp, 0, (size_t)c * cnt * sizeof(double), h->st);
    rc = stage_sweep_multi(*h, c, d_mu + i, tmp);
    if (rc == OK && cudaMemcpyAsync(w_out + (size_t)i * cnt, tmp, (size_t)c * cnt * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess)
      rc = fail(ECUDA, "sweep copy failed");
    cudaStreamSynchronize(h->st);
  }
  cudaFree(tmp);
  cudaFree(d_mu);
  return rc;
}

/* Sweep without leaving the device: filters for every mu into a caller-owned DEVICE buffer (n_mu, 2, V, n) (may be
 * NULL) and the eigen-basis figures of merit (n_mu, 2, V, 3) = {dark energy, bright energy, w . r_B} per rank to the
 * host (may be NULL).  The full-rank sweep of BASELINE cfg-4 is 2 GB of filters per block: they stay in HBM. */
static int sweep_scratch(Handle& h, size_t n_mu, size_t n_metrics) {
  if (n_mu > h.sw_mu_cap) {
    if (h.sw_mu) cudaFree(h.sw_mu);
    h.sw_mu = nullptr; h.sw_mu_cap = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&h.sw_mu, std::max<size_t>(n_mu, 64) * sizeof(double)));
    h.sw_mu_cap = std::max<size_t>(n_mu, 64);
  }
  if (n_metrics > h.sw_m_cap) {
    if (h.sw_m) cudaFree(h.sw_m);
    h.sw_m = nullptr; h.sw_m_cap = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&h.sw_m, n_metrics * sizeof(double)));
    h.sw_m_cap = n_metrics;
  }
  return OK;
}

int apv_sweep_device(apv_handle* h, int n_mu, const double* mu, void* d_w_out, double* metrics_out) {
  if (!h || !mu || n_mu < 1) return fail(EINVAL_, "bad argument");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const size_t mc = (size_t)n_mu * 2 * D.V * 3;
  APV_TRY(sweep_scratch(*h, (size_t)n_mu, metrics_out ? mc : 0));
  double* d_mu = h->sw_mu;
  int rc = OK;
  if (cudaMemcpyAsync(d_mu, mu, n_mu * sizeof(double), cudaMemcpyHostToDevice, h->st) != cudaSuccess) rc = fail(ECUDA, "mu copy failed");
  if (rc == OK && d_w_out) {
    if (!(D.runA && D.runB)) cudaMemsetAsync(d_w_out, 0, (size_t)n_mu * 2 * D.V * D.n * sizeof(double), h->st);
    rc = stage_sweep_multi(*h, n_mu, d_mu, (double*)d_w_out);
  }
  if (rc == OK && metrics_out) {
    cudaMemsetAsync(h->sw_m, 0, mc * sizeof(double), h->st);
    rc = stage_sweep_metrics(*h, n_mu, d_mu, h->sw_m);
    if (rc == OK && cudaMemcpyAsync(metrics_out, h->sw_m, mc * sizeof(double), cudaMemcpyDeviceToHost, h->st) != cudaSuccess)
      rc = fail(ECUDA, "metrics copy failed");
  }
  if (cudaStreamSynchronize(h->st) != cudaSuccess && rc == OK) rc = fail(ECUDA, "sweep failed: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

int apv_eval_zone(apv_handle* h, int zone, int n_samples, const double* feeds, const double* signal, double* out3) {
  if (!h || !feeds || !signal || !out3) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
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
  DevGuard dg(h->device);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st_front));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st_copy));
  for (auto& sx : h->st_backx) APV_CUDA_TRY(cudaStreamSynchronize(sx));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_stage_times(apv_handle* h, float* ms7) {
  if (!h || !ms7) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st_front));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  // front half: ev[0] .. ev[3]; back half: ev[7], ev[4] .. ev[6] (they overlap other blocks in a multi-block call)
  const int from[6] = {0, 1, 2, 7, 4, 5}, to[6] = {1, 2, 3, 4, 5, 6};
  ms7[6] = 0.f;
  for (int i = 0; i < 6; ++i) {
    ms7[i] = 0.f;
    cudaEventElapsedTime(&ms7[i], h->ev[from[i]], h->ev[to[i]]);
    ms7[6] += ms7[i];
  }
  return OK;
}

int apv_jdiag_phase_times(apv_handle* h, float* ms6) {
  if (!h || !ms6) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  for (int i = 0; i < 6; ++i) ms6[i] = 0.f;
  if (h->nz == 0) return OK;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms6[i], h->jd.ev[i], h->jd.ev[i + 1]);
  return OK;
}

int apv_kernel_times(apv_handle* h, float* ms4) {
  if (!h || !ms4) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  for (int i = 0; i < 4; ++i) ms4[i] = 0.f;
  APV_CUDA_TRY(cudaStreamSynchronize(h->st_front));
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
  DevGuard dg(h->device);
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  APV_CUDA_TRY(cudaEventRecord(h->ev_timer[0], h->st));
  return OK;
}

int apv_timer_stop(apv_handle* h, float* ms) {
  if (!h || !ms) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
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

int apv_bench_gemm_shape(int M, int N, int K, int batch, int transB, int tri, int mirror, int bn, double beta,
                         int nrep, float* ms) {
  if (M < 1 || N < 1 || K < 1 || batch < 1 || nrep < 1 || !ms) return fail(EINVAL_, "bad argument");
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  const size_t lda = (size_t)round_up(K, 2), ldb = transB ? lda : (size_t)round_up(N, 2), ldc = (size_t)round_up(N, 2);
  const size_t ca = (size_t)M * lda, cb = transB ? (size_t)N * ldb : (size_t)K * ldb, cc = (size_t)M * ldc;
  APV_CUDA_TRY(cudaMalloc((void**)&dA, batch * ca * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dB, batch * cb * sizeof(double)));
  APV_CUDA_TRY(cudaMalloc((void**)&dC, batch * cc * sizeof(double)));
  fill_kernel<<<256, 256>>>(dA, batch * ca, 1.0 / K);
  fill_kernel<<<256, 256>>>(dB, batch * cb, 0.5);
  fill_kernel<<<256, 256>>>(dC, batch * cc, 0.0);
  GemmArgs g{};
  g.A = dA; g.B = dB; g.C = dC; g.M = M; g.N = N; g.K = K; g.lda = (int)lda; g.ldb = (int)ldb; g.ldc = (int)ldc;
  g.strideA = (long long)ca; g.strideB = (long long)cb; g.strideC = (long long)cc;
  g.alpha = beta != 0.0 ? 1e-3 : 1.0; g.beta = beta; g.batch = batch; g.transB = transB; g.tri = tri; g.mirror = mirror; g.bn = bn;
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
