// Internal structures of the AP-VAST B200 engine: device state of one handle and the stage launchers.
#pragma once
#include "../../include/apvast_b200.h"
#include "common.cuh"

namespace apv {

constexpr int STATS_KC = 128;   // K chunk of the statistics SYRK; fixes the padding of the packed s' buffer

// Derived sizes.  Symbols as in SURVEY.md: Nb block, H hop, K rir length, L srcs, M mics, J taps,
// N statistics length, V ranks, n = L*J, P = N-J columns of the data matrix, F = Nb/2+1 bins.
struct Dims {
  int Nb, H, K, L, M, J, N, V, d, refA, refB, runA, runB, F, n, ldn, P, Ns, LX;
  int clean;       // 1: MATLAB data matrix (N-J+1 columns, no skipped sample)
};

// ---------------------------------------------------------------------------------------------
// Joint diagonalisation workspace (jdiag.cu); batched over `nz` zone problems of equal size.
struct JdiagWs {
  int n = 0, ldn = 0, V = 0, nz = 0, nb = 0, nbt = 0, eig_mode = 0;
  double* Lm = nullptr;     // [nz][n][ldn]  R_D + reg I -> Cholesky factor (lower)
  double* Cm = nullptr;     // [nz][n][ldn]  R_B -> C = L^-1 R_B L^-T -> (tridiagonalised in place)
  double* Tm = nullptr;     // [nz][n][ldn]  scratch (transpose)
  double* VH = nullptr;     // [nz][n][ldn]  Householder vectors, row j = v_j
  double* Dinv = nullptr;   // [nz][n/nb][nb][nb] inverses of the Cholesky diagonal blocks
  double* SBinv = nullptr;  // [nz][n/256][256][256] inverses of the diagonal super-blocks (triangular solves)
  double* Z1 = nullptr;     // [nz][n][2 nbt]  [V | W] panel
  double* Z2 = nullptr;     // [nz][n][2 nbt]  [W | V] panel
  double* tau = nullptr;    // [nz][n]
  double* dd = nullptr;     // [nz][n] diagonal of T
  double* ee = nullptr;     // [nz][n] off-diagonal of T
  double* colbuf = nullptr; // [nz][n]
  double* ybuf = nullptr;   // [nz][n]
  double* wbuf = nullptr;   // [nz][n]
  double* tdws = nullptr;   // scratch of the tridiagonalisation panel kernel (partials, tile partial vectors)
  double* Tf = nullptr;     // [nz][n/8][8][8] compact-WY T factors of the reflector blocks
  double* lam = nullptr;    // [nz][V]   top-V eigenvalues, descending
  double* shift = nullptr;  // [nz][V]   perturbed shifts for inverse iteration
  double* iv = nullptr;     // [nz][6][n][Vp] inverse-iteration work (interleaved over vectors)
  double* Zt = nullptr;     // [nz][V][n] eigenvectors of T -> of C -> joint eigenvectors U (row v)
  int* info = nullptr;      // [nz][4]: [0] first non-positive pivot (1-based, 0 = ok) [1] eig flags
  double* dcw = nullptr;    // divide-and-conquer scratch (dc.cu; allocated on first use)
  size_t dcw_count = 0;
  double* q1agg = nullptr;  // aggregated block reflectors + scratch of the GEMM back-transformation (allocated on first use)
  size_t q1agg_count = 0;
  double* ts2 = nullptr;    // two-stage tridiagonalisation scratch (band.cu): V panel, Y slices, X, S partials, T, band, flags
  cudaEvent_t ev2[4] = {};  // two-stage: end of stage 1 (dense -> band), end of stage 2 (band -> tridiagonal), look-ahead fork / join
  cudaStream_t st2 = nullptr;   // high-priority side stream of the look-ahead panel factorisation
  int Vp = 0;
  size_t bytes = 0;
  cudaEvent_t ev[8] = {};   // phase boundaries: prep | chol | reduce | tridiag | eig | backtransform | solve
  cudaEvent_t* pev = nullptr;   // [2 * npanel] events around every td_panel_kernel launch (roofline timing)
  int npanel = 0;
  bool last_two_stage = false;  // the last jdiag_run used the two-stage tridiagonalisation (band.cu)
  bool last_panels = false;     // ... the one-stage panel kernels (pev recorded)
};
int jdiag_alloc(JdiagWs& ws, int n, int V, int nz, int eig_mode);
void jdiag_free(JdiagWs& ws);
// bright[z], dark[z]: device n x n matrices with leading dimension ld_in.
// tridiag.cu: Cm -> (dd, ee, VH, tau); uses Z1/Z2/colbuf/tdws.  Adds its kernel launches to *launches.
int tridiag_run(JdiagWs& ws, cudaStream_t st, int* launches);
size_t tridiag_scratch_doubles(int n, int nz);
// band.cu (eig_mode 3): Cm -> band (stage-1 reflectors in VH / tau) -> (dd, ee) (stage-2 reflectors in Tm);
// twostage_apply_q2 maps the eigenvectors of T in the inverse-iteration workspace to those of the band matrix.
size_t twostage_scratch_bytes(int n, int nz, int nsplit_max);
int twostage_nsplit_max();
int twostage_run(JdiagWs& ws, cudaStream_t st, int* launches);
int twostage_apply_q2(JdiagWs& ws, cudaStream_t st, int* launches);
int twostage_apply_q1(JdiagWs& ws, cudaStream_t st, int* launches);   // then by the stage-1 block reflectors -> Zt
int twostage_apply_q1_gemm(JdiagWs& ws, cudaStream_t st, int* launches);
// dc.cu: all n eigenpairs of the tridiagonal matrix (dd, ee) by divide and conquer -> lam (descending), iv slot 4
bool dc_supported(int n);
int dc_run(JdiagWs& ws, cudaStream_t st, int* launches);   // many vectors: DMMA GEMMs, in place in iv slot 4
// regv != nullptr: per-zone diagonal loading read from device memory instead of `reg`.
int jdiag_run(JdiagWs& ws, const double* const bright[2], const double* const dark[2], int ld_in, double reg,
              cudaStream_t st, int* launches, const double* regv = nullptr);

// ---------------------------------------------------------------------------------------------
struct Handle {
  apv_config cfg;
  Dims D;
  cudaStream_t st = nullptr;       // main stream (highest priority): S5-S7, and S1-S4 too in the per-block call
  cudaStream_t st_front = nullptr; // low-priority stream: S1-S4 of block t+1 while S5-S7 of block t run on `st`
  cudaStream_t st_copy = nullptr;  // D2H of the rendered outputs of the multi-block call
  cudaEvent_t ev[8] = {};          // [0] front start [1] S1 [2] S3 [3] S4 | [7] back start [4] S5 [5] S6 [6] S7
  cudaEvent_t ev_syrk[2] = {};   // around the statistics SYRK kernel
  cudaEvent_t ev_timer[2] = {};  // apv_timer_start / apv_timer_stop
  int device = 0;
  // constants
  double* rirT = nullptr;    // [zone 2][M][L][K]
  double* rirTT = nullptr;   // [zone 2][M][K]     delayed target RIRs (apvast.py:102-112)
  double* win = nullptr;     // [Nb]
  double2* tw = nullptr;     // [Nb]  exp(-2 pi i k / Nb)
  int nrad = 0;
  int rad[32];
  double* G2 = nullptr;      // [C][F] masking model table
  int nchan = 0;
  double Cs = 0, Ca = 0, Leff = 1;
  // state
  double* xin = nullptr;     // [2][LX]
  double* Q = nullptr;       // [4][M][L][Nb]
  double* QT = nullptr;      // [2][M][Nb]
  double* O = nullptr;       // [4][M][L][Nb]
  double* OT = nullptr;      // [2][M][Nb]
  double* S = nullptr;       // [4][M][L][N]
  double* ST = nullptr;      // [2][M][N]
  double* Sp = nullptr;      // [4][M][L][Ns]   s' = delete(S, J), zero padded
  double* seed = nullptr;    // [4][L][L][J]    first-row correlations (stats_mode 2)
  double* norms = nullptr;   // [4] spectral norms of the statistics (loading_mode 1) + power-iteration scratch
  double* pvec = nullptr;    // [2][4][n] power-iteration vectors
  double* Pbuf = nullptr;    // [syrk_slots][16 slices][128][128] ring of per-microphone partial tiles (DMMA SYRK, tree-summed)
  int* syrk_cnt = nullptr;   // [4][lower tiles] arrival counters of the fused reduction (self-resetting) + [syrk_slots] ring state
  int syrk_group = 16;       // microphones per SYRK launch (at most 16)
  int syrk_slots = 16;       // ring slots (16 x 2 MB: stays in L2; 12 and 24 measure the same, 48 is 1.7 ms slower)
  double* Wg = nullptr;      // [2][M][F]
  double* tframe = nullptr;  // [2][M][Nb]
  double2* tspec = nullptr;  // [2][M][Nb]      target spectra (split call)
  double* G = nullptr;       // [2][V][L][Nb]
  double* Gt = nullptr;      // [2][Nb]
  // results.  R, rvec and xw exist in two slots so that the statistics of block t+1 can be formed while block t is
  // still being diagonalised and rendered (apv_process_blocks / apv_range_run); R, rvec, xw point at the slot of the
  // block whose S5-S7 ran last.
  static constexpr int MAXDEPTH = 4;        // back halves in flight
  static constexpr int NSLOT = MAXDEPTH + 1; // statistics slots: every back half in flight holds one, a front half fills one
  int nslot = 2;                             // slots in use = depth + 1 (allocated on demand beyond the first two)
  double* Rslot[NSLOT] = {};                 // [4][n][ldn]
  double* rvslot[NSLOT] = {};                // [2][n]
  double* xwslot[NSLOT] = {};                // [2][Nb]  window * input block (what S7 filters, apvast.py:430-431)
  cudaEvent_t ev_ready[NSLOT] = {};          // front of the slot finished (statistics ready)
  cudaEvent_t ev_free[NSLOT] = {};           // back of the slot finished (slot may be overwritten)
  cudaEvent_t ev_order = nullptr;            // S6 + S7 of the previous block finished (the overlap buffers G are sequential state)
  cudaEvent_t ev_join = nullptr;             // hand-over between the main stream and the others at the ends of a call
  cudaStream_t st_backx[MAXDEPTH - 1] = {};  // further back-half streams (depth 2..4); the first back-half stream is `st`
  int depth = 1;                             // back halves (S5-S7) in flight in a multi-block call: 1..4
  int last_ws = 0;                           // joint-diagonalisation workspace the last back half used
  double* R = nullptr;       // [4][n][ldn]
  double* rvec = nullptr;    // [2][n]
  double* xw = nullptr;      // [2][Nb]
  double* regv = nullptr;    // [2] per-zone diagonal loading when EXPERIMENTAL_REGULARIZATION is off (apvast.py:25-27)
  double* lam = nullptr;     // [2][V]
  double* U = nullptr;       // [2][V][n]
  double* W = nullptr;       // [2][V][n]
  // staging
  double* d_in = nullptr;    // [2][H]
  double* d_out = nullptr;   // [2][V][H][L]
  double* d_out_t = nullptr; // [2][H]
  double* h_pin = nullptr;   // pinned host staging
  size_t h_pin_count = 0;
  double* sw_mu = nullptr;   // persistent scratch of the mu sweep (cudaMalloc / cudaFree per call cost up to 0.4 s)
  double* sw_m = nullptr;
  size_t sw_mu_cap = 0, sw_m_cap = 0;
  double* home_W = nullptr;  // the buffers W / d_out / d_out_t point at outside a multi-block call
  double* home_out = nullptr;
  double* home_out_t = nullptr;
  // multi-block call: ring of rendered blocks in HBM + pinned host ring for the asynchronous D2H
  double* ring = nullptr;    // [ring_cap] x (out 2 V H L | out_t 2 H | W 2 V n)
  int ring_cap = 0;
  double* ring_pin = nullptr;
  int ring_pin_cap = 0;
  int* ring_info = nullptr;  // pinned [ring_pin_cap][8]
  cudaEvent_t ev_rend[8] = {};   // block rendered into its ring slot
  cudaEvent_t ev_d2h[8] = {};    // ring slot copied to the pinned ring
  int pipeline = 1;          // 0: S1-S7 of consecutive blocks strictly in order on one stream
  // block-range sharding (comm.cu)
  double* rg_out = nullptr;  // [rg_cap][2][V][H][L] rendered outputs of the owned blocks
  double* rg_w = nullptr;    // [rg_cap][2][V][n]    filters of the owned blocks
  int* rg_info = nullptr;    // [rg_cap][8] jdiag status of the owned blocks
  int rg_cap = 0, rg_owned = 0;
  double* rg_sig = nullptr;  // [2][halo + owned][H] hops of a range copied from the host
  size_t rg_sig_cap = 0;
  int* gat_info = nullptr;   // root: [total blocks][8]
  double* rg_tail_send = nullptr;  // [2][V][L][Nb-H]
  double* rg_tail_recv = nullptr;
  double* gat_out = nullptr; // root: [total blocks][2][V][H][L]
  double* gat_w = nullptr;   // root: [total blocks][2][V][n]
  int gat_cap = 0;
  void* comm = nullptr;      // ncclComm_t
  int comm_rank = 0, comm_size = 1;
  JdiagWs jd;
  JdiagWs jdx[3];            // further workspaces: up to four joint diagonalisations in flight (allocated on demand)
  int nz = 0;
  int zones[2] = {0, 1};
  float stage_ms[7] = {};
  int launches = 0;
  int launches_front = 0;
  cudaEvent_t* dbg_ev = nullptr;   // APV_PIPE_DEBUG: [dbg_cap][4] front start / front end / back start / back end per block
  int dbg_cap = 0;    // launches of the front half of the block whose back half runs next (multi-block calls)
  bool began = false;
};

// stage launchers (each returns status; `launches` is incremented per kernel launch)
int stage_fir(Handle& h, const double* d_inA, const double* d_inB);           // S1 (fir_wola.cu); also fills h.xw
int stage_targets(Handle& h, bool compute_gain);                              // S2 + S2b
int stage_weighted(Handle& h);                                                // S3
int stage_stats(Handle& h);                                                   // S4 (stats.cu)
int stage_loading(Handle& h);                                                 // MATLAB diagonal loading (stats.cu)
int stage_sweep(Handle& h, double mu, double* W_out);                         // S6 (render.cu)
int stage_sweep_multi(Handle& h, int n_mu, const double* d_mu, double* W_out);     // list of mu, one launch
int stage_sweep_metrics(Handle& h, int n_mu, const double* d_mu, double* d_out);   // eigen-basis metrics of the sweep
int stage_render(Handle& h);                                                  // a2 + S7
int eval_zone(Handle& h, int zone, int T, const double* feeds, const double* signal, double* out3);  // metrics.cu
int stage_spectral_norms(Handle& h);                                          // |R_D|_2 per zone -> regv (stats.cu)
int fft_plan(int n, int* rad, int* nrad);
// engine.cu internals used by comm.cu
int run_front(Handle& h, const double* d_inA, const double* d_inB, bool skip_s1, bool state_only);
int run_back(Handle& h, cudaEvent_t order = nullptr);
int fail(int code, const char* fmt, ...);
struct DevGuard {   // entry points run on the handle's device whatever the caller's current device is
  int prev = -1, dev = -1;
  explicit DevGuard(int d) : dev(d) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev != dev && prev >= 0) cudaSetDevice(prev);
  }
};
struct BlockSink {   // device destinations of one block's results (nullptr = the handle's own buffers)
  double* out;     // [2][V][H][L]
  double* out_t;   // [2][H]
  double* W;       // [2][V][n]
  int* info;       // [8] status of the joint diagonalisation
};
bool pipelined(const Handle& h);
int enqueue_front(Handle& h, long b, const double* d_inA, const double* d_inB, bool state_only);
int enqueue_back(Handle& h, long b, const BlockSink& sink, cudaEvent_t done = nullptr);
int ensure_depth(Handle& h, int depth);
int leave_multiblock(Handle& h);
int status_from_info(const Handle& h, const int* info, long block);
int range_alloc(Handle& h, int max_owned, int total_on_root);
void range_free(Handle& h);
int fft_util(int n, int inverse, const double* in_ri, double* out_ri);        // test hook

}  // namespace apv

struct apv_handle : public apv::Handle {};
