/*
 * apvast_b200.h -- C-ABI of the B200-native AP-VAST block engine.
 *
 * Drop-in boundary for the hot path of macoustics/ap-vast-unofficial: the Python class `apvast`
 * (reference Python/apvast.py:39) -- constructor (:40-151) and per-block call
 * `process_input_buffers(input_A, input_B)` (:153-165) -- and the `jdiag` it calls (:20-36).
 * The reference has no FFI of its own (it is pure NumPy/SciPy); these entry points are what a
 * ctypes binding for that path binds (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; all arithmetic is IEEE float64; all host arrays are C-order.
 *   - every function returns 0 (APV_OK) or a positive status; apv_last_error() gives the message.
 *   - one handle = one ordered stream of hops (like one reference object); calls on a handle are
 *     serialised by the caller.  One CUDA stream per handle.
 *   - no cuBLAS / cuSOLVER / cuFFT symbols are linked; there is no CPU fallback.
 */
#ifndef APVAST_B200_H
#define APVAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum apv_status {
  APV_OK = 0,
  APV_EINVAL = 1,  /* bad argument / size (reference raises RuntimeError, apvast.py:86-90,154-155) */
  APV_ENOTPD = 2,  /* R_D + reg*I not positive definite (reference: numpy LinAlgError, apvast.py:21-24) */
  APV_ECUDA = 3,   /* CUDA runtime error */
  APV_ENOMEM = 4,
  APV_ENOCONV = 5, /* eigen-solver did not converge */
  APV_ENCCL = 6    /* NCCL error (block-range sharding) */
};

/* Constructor parameters: reference apvast.__init__ (apvast.py:40-56) + module flags (:6-7). */
typedef struct apv_config {
  int32_t block_size;        /* Nb, even                          (ctor block_size, :41)            */
  int32_t hop_size;          /* H, 0 -> Nb/2                      (ctor hop_size, :51,93)           */
  int32_t rir_length;        /* K = rir.shape[0]                  (:97)                             */
  int32_t n_srcs;            /* L = rir.shape[1]                  (:98)                             */
  int32_t n_mics;            /* M = rir.shape[2]                  (:99)                             */
  int32_t filter_length;     /* J                                 (:44)                             */
  int32_t stats_length;      /* N = statistics_buffer_length      (:50)                             */
  int32_t n_eig;             /* V = number_of_eigenvectors        (:48)                             */
  int32_t modeling_delay;    /* d                                 (:45)                             */
  int32_t ref_A, ref_B;      /* reference_index_A/B, 0-based      (:46-47)                          */
  int32_t run_A, run_B;      /* (:53-54)                                                            */
  int32_t perceptual;        /* 0: W == 1 (:326-327); 1: on-device masking model (apv_set_gain_table);
                                2: weighting spectra supplied by the host per block (apv_begin_block, APV_T_WEIGHT);
                                3: on-device model, split call: apv_begin_block runs S1 + S2, the caller may then
                                   exchange weighting curves between handles (apv_copy_weights), apv_finish_block
                                   runs S3..S7 */
  int32_t normalize_gains;   /* EXPERIMENTAL_NORMALIZE_GAINS (:6,322-324)                           */
  int32_t eig_mode;          /* 0 auto (2 for n <= 48, 3 for n >= 1024, else 1);
                                1 one-stage Householder tridiagonalisation + bisection + inverse
                                  iteration (top-V);
                                2 cyclic Jacobi (small n, full spectrum);
                                3 two-stage tridiagonalisation: dense -> band (DMMA, cluster QR)
                                  -> tridiagonal (bulge chasing), then as 1                         */
  int32_t stats_mode;        /* 0/1 FP64 tensor-core SYRK with implicit Toeplitz operand (default);
                                2 structured evaluation: first-row correlations + double-double
                                  diagonal recurrence (~J/2 x fewer flops, opt-in)                  */
  int32_t device;            /* CUDA device ordinal, -1 = current                                   */
  double mu;                 /* (:49)                                                               */
  double reg;                /* absolute diagonal loading inside jdiag, reference 1e-7 (:22-24)     */
  double sampling_rate;      /* (:52)                                                               */
  /* MATLAB-flavoured variant switches (Matlab/ControlMethods/apVast.m; all 0 = the Python reference): */
  int32_t toeplitz_clean;    /* 1: N-J+1 columns, no skipped sample (apVast.m:420-422)              */
  int32_t normalize_stats;   /* 1: R, r divided by (N-J+1)*M (apVast.m:448-456)                     */
  int32_t loading_mode;      /* 0: dark += reg*I inside jdiag (apvast.py:22-24);
                                1: bright += bright_load*|R_B|_2*I, dark += dark_load*|R_D|_2*I, applied
                                   to the stored statistics, no further regularisation (apVast.m:552-569) */
  int32_t target_ref_per_zone; /* 1: target of zone B uses reference_index_B (apVast.m:597-600)     */
  double bright_load;        /* brightCondLimit, MATLAB 1e-8 (apVast.m:560)                         */
  double dark_load;          /* darkCondLimit, MATLAB 5e-3 (apVast.m:559)                           */
  /* more than two zones by composition (zones.py): the microphones m >= active_mics_A of zone A are silent padding
     (zero RIRs) and are skipped by the statistics; 0 = all n_mics are real */
  int32_t active_mics_A;
  int32_t reg_relative;      /* EXPERIMENTAL_REGULARIZATION off (apvast.py:25-27): dark += 1e-8 |R_D|_2 I inside
                                jdiag instead of the absolute `reg` (spectral norm by power iteration)        */
} apv_config;

typedef struct apv_handle apv_handle;

/* Tensor ids for apv_get / apv_set.  Device layouts are channel-major, time-contiguous;
 * the Python host transposes to the reference's attribute shapes. */
enum apv_tensor {
  APV_T_W = 0,            /* [zone 2][V][n]                 filters w_A, w_B (apvast.py:393-414)            */
  APV_T_LAMBDA = 1,       /* [zone 2][V]                    top-V eigenvalues, descending (:385-387)        */
  APV_T_U = 2,            /* [zone 2][V][n]                 top-V joint eigenvectors, row v = U[:, v]       */
  APV_T_R = 3,            /* [path 4][n][n]                 R_A_to_A, R_A_to_B, R_B_to_A, R_B_to_B (:368-372)*/
  APV_T_RVEC = 4,         /* [zone 2][n]                    r_A, r_B (:373-376)                             */
  APV_T_WEIGHT = 5,       /* [zone 2][M][F]                 real weighting gains W_A, W_B (:315-327)        */
  APV_T_RESP = 6,         /* [path 4][M][L][Nb]             loudspeaker_response_*_buffer (:124-127)        */
  APV_T_RESP_T = 7,       /* [zone 2][M][Nb]                loudspeaker_target_response_*_buffer (:128-129) */
  APV_T_OLA = 8,          /* [path 4][M][L][Nb]             ..._weighted_response_*_overlap_buffer (:132-135)*/
  APV_T_OLA_T = 9,        /* [zone 2][M][Nb]                ..._weighted_target_*_overlap_buffer (:136-137) */
  APV_T_STATS = 10,       /* [path 4][M][L][N]              ..._weighted_response_*_buffer (:140-143)       */
  APV_T_STATS_T = 11,     /* [zone 2][M][N]                 ..._weighted_target_response_*_buffer (:144-145)*/
  APV_T_OUT_OLA = 12,     /* [zone 2][V][L][Nb]             output_A/B_overlap_buffer (:148-149)            */
  APV_T_OUT_OLA_T = 13,   /* [zone 2][Nb]                   output_*_t_overlap_buffer, reference loudspeaker
                                                            column (all other columns are zero, :150-151)   */
  APV_T_INPUT = 14,       /* [signal 2][LX]                 most recent input samples, newest last;
                                                            LX = max(K-1+H, Nb) (input_*_block + FIR history) */
  APV_T_TARGET_FRAME = 15 /* [zone 2][M][Nb]                irfft of the windowed target spectra: what the
                                                            reference passes to model.gain (:318-319)       */
};

/* Element count of a tensor (0 if unknown id). */
size_t apv_tensor_size(const apv_handle* h, int tensor_id);

/* Constructor.  rir_A/rir_B: (K, L, M) C-order (apvast.py:42-43).  init_resp: the six 1e-3*randn start
 * buffers in the reference draw order (:124-129): 4 x (Nb, L, M) then 2 x (Nb, M), C-order, concatenated;
 * NULL = zeros. */
int apv_create(const apv_config* cfg, const double* rir_A, const double* rir_B, const double* init_resp,
               apv_handle** out);
void apv_destroy(apv_handle* h);

/* One filter update = reference process_input_buffers (apvast.py:153-165).  in_A/in_B: H samples (host).
 * out_A/out_B: (V, H, L) C-order host buffers or NULL; out_A_t/out_B_t: (H, L) C-order host buffers or
 * NULL (the reference returns V identical copies of these, :418,422,467-475). */
int apv_process_block(apv_handle* h, const double* in_A, const double* in_B, double* out_A, double* out_B,
                      double* out_A_t, double* out_B_t);

/* Throughput path: nblocks consecutive hops in one call; results identical to nblocks per-block calls (the caller of
 * the reference is a plain per-hop loop, make_python_test.m:44-54).  The hops are copied to the device once; S1-S4 of
 * block t+1 (which depend on the streaming state only, apvast.py:329-364) run on a second, low-priority stream while
 * S5-S7 of block t occupy the main stream; rendered blocks leave through a ring in HBM, a copy stream and a pinned host
 * ring, one block behind.  in_*: (nblocks, H); out_A/out_B: (nblocks, V, H, L) or NULL; out_*_t: (nblocks, H, L) or
 * NULL; w_out: (nblocks, 2, V, n) or NULL (per-block filters). */
int apv_process_blocks(apv_handle* h, int nblocks, const double* in_A, const double* in_B, double* out_A,
                       double* out_B, double* out_A_t, double* out_B_t, double* w_out);

/* Same as apv_process_block but inputs/outputs are DEVICE pointers and nothing is copied or synchronised
 * (used to time the kernels with inputs resident in HBM). */
int apv_process_block_device(apv_handle* h, const double* d_in_A, const double* d_in_B);

/* Split call for perceptual == 2 (host gain model, e.g. libdetectability, apvast.py:318-319):
 * begin = S1 + target spectra; read APV_T_TARGET_FRAME; set APV_T_WEIGHT; finish = S2..S7. */
int apv_begin_block(apv_handle* h, const double* in_A, const double* in_B);
int apv_finish_block(apv_handle* h, double* out_A, double* out_B, double* out_A_t, double* out_B_t);

/* Warm-up only: S1-S3 (state update) without statistics / filters / rendering (multi-GPU halo). */
int apv_advance_state(apv_handle* h, const double* in_A, const double* in_B);

/* 0: consecutive blocks of a multi-block call strictly in order on one stream; 1: S1-S4 of block t+1 start together
 * with S5-S7 of block t; 2 (default, also env APV_PIPELINE): they start when block t reaches its bulge chasing, the
 * point from which S5 leaves most SMs idle.  Results are bit-identical in all three. */
int apv_set_pipeline(apv_handle* h, int on);
/* Back halves (S5-S7) in flight in a multi-block call, 1..4: the joint diagonalisations of consecutive blocks run side
 * by side on their own streams and workspaces while S6/S7 stay in block order.  Default 2 (4 for n < 2048, where S5 is
 * latency-bound throughout); the extra workspaces are allocated by the first multi-block call; env APV_DEPTH
 * overrides.  Results are bit-identical. */
int apv_set_depth(apv_handle* h, int depth);
/* Diagnostic timeline of a multi-block call: arm with ms == NULL, run, read back (nblocks, 4) milliseconds
 * (front start, front end, back start, back end per block, relative to the first event). */
int apv_debug_timeline(apv_handle* h, int nblocks, float* ms);
/* EXPERIMENTAL_REGULARIZATION (apvast.py:7,22-27), read by the reference at call time: 0 = absolute reg (default),
 * 1 = 1e-8 |R_D|_2 (spectral norm by power iteration on the device). */
int apv_set_reg_mode(apv_handle* h, int relative);
/* perceptual == 3 (more than two zones, zones.py): after apv_begin_block every handle holds the weighting curves of
 * its own bright zone (zone 0); the dark microphones of handle `dst` that belong to the bright zone of handle `src`
 * take src's curves (each dark microphone is weighted from its own zone's target: the generalisation of
 * apvast.py:259-262,318-319).  Device to device, same GPU. */
int apv_copy_weights(apv_handle* dst, int dst_zone, int dst_mic0, apv_handle* src, int src_zone, int src_mic0, int n_mics);

/* ---- block-range sharding over the GPUs of a box (one process per GPU, NCCL over NVLink bound at run
 * time with dlopen -- the library links no NCCL symbol).  The reference is one ordered stream of hops
 * (apvast.py:153-165); its streaming state has finite memory (apvast.py:115-151), so rank g replays S1-S3 over a halo
 * and owns a contiguous block range; the output overlap-add tail (apvast.py:455-465) is the only data that crosses a
 * range boundary. */
int apv_comm_unique_id(void* id128);                                        /* rank 0: ncclGetUniqueId, 128 bytes */
int apv_comm_init(apv_handle* h, int rank, int nranks, const void* id128);  /* nranks == 1: no communicator needed */
int apv_comm_destroy(apv_handle* h);
int apv_nccl_version(int* version);
/* device buffers for ranges of up to max_halo + max_owned blocks; total_on_root = blocks the gathering rank receives */
int apv_range_reserve(apv_handle* h, int max_halo, int max_owned, int total_on_root);
/* n_halo state-only blocks, then n_owned full blocks (pipelined); in_*: (n_halo + n_owned) * H samples, host or device
 * pointers.  Asynchronous: returns once everything is enqueued; outputs and filters of the owned blocks stay in HBM. */
int apv_range_run(apv_handle* h, int n_halo, int n_owned, const double* in_A, const double* in_B, int inputs_on_device);
/* overlap-add halo: G[:, :, H:] of rank g -> rank g+1 (ncclSend/ncclRecv on the handle's stream), added to the first
 * Nb/H - 1 owned output blocks there.  Every range but the last must hold at least Nb/H - 1 blocks. */
int apv_range_exchange_halo(apv_handle* h);
/* outputs (counts[r] blocks of (2, V, H, L)) and filters ((2, V, n)) of every rank, in rank order, into the HBM of
 * `root` and from there to out_host / w_host (root only, may be NULL; pinned memory recommended: apv_alloc_pinned).
 * Synchronises; returns the first joint-diagonalisation failure of any block. */
int apv_range_gather(apv_handle* h, int root, const int* counts, double* out_host, double* w_host);
int apv_range_device_ptrs(apv_handle* h, void** gathered_out, void** gathered_w, void** own_out, void** own_w);
/* test hooks: the packed tail this handle would send / adding a tail without a communicator */
int apv_range_tail_get(apv_handle* h, double* tail_host);
int apv_range_tail_add(apv_handle* h, const double* tail_host);
void* apv_alloc_pinned(size_t bytes);
void apv_free_pinned(void* p);

int apv_get(apv_handle* h, int tensor_id, double* dst, size_t count);
int apv_set(apv_handle* h, int tensor_id, const double* src, size_t count);
int apv_set_mu(apv_handle* h, double mu);

/* Perceptual masking model tables (perceptual == 1): G2[c][f] = (outer/middle-ear x gammatone)^2,
 * C channels x F = Nb/2+1 bins, and the calibration constants (perceptualModel.m:30-139). */
int apv_set_gain_table(apv_handle* h, int n_channels, const double* G2, double Cs, double Ca, double Leff);

/* mu x V trade-off sweep (BASELINE cfg-4): filters for n_mu values of mu from ONE joint diagonalisation.
 * w_out: (n_mu, 2, V, n) host buffer. */
int apv_sweep(apv_handle* h, int n_mu, const double* mu, double* w_out);
/* The same sweep without leaving the device: d_w_out = DEVICE buffer (n_mu, 2, V, n) or NULL; metrics_out = host
 * (n_mu, 2, V, 3) or NULL: per rank v the eigen-basis figures of merit of w[v] -- dark energy w'(R_D + reg I)w = sum a_i^2,
 * bright energy w'R_B w = sum lambda_i a_i^2 and w'r_B = sum a_i c_i (a_i = c_i / (lambda_i + mu), c = U'r_B); they follow
 * from U'(R_D + reg I)U = I, U'R_B U = Lambda (jdiag.m:33-35) and pick the operating point of the trade-off. */
int apv_sweep_device(apv_handle* h, int n_mu, const double* mu, void* d_w_out, double* metrics_out);

/* Evaluation of rendered loudspeaker feeds (callers' side of the path: Matlab/ControlMethods/predictPressure.m:12-17,
 * Matlab/main.m:120-130).  feeds: (n_samples, L) host, signal: (n_samples) programme signal of the zone.
 * out3: [0] acoustic contrast at the control microphones [dB], [1] NMSE between target and bright-zone pressure
 * (mean over microphones), [2] 10 log10(NMSE) = normalised signal distortion [dB]. */
int apv_eval_zone(apv_handle* h, int zone, int n_samples, const double* feeds, const double* signal, double* out3);

/* Device pointer of a tensor (for torch.distributed / NCCL plumbing on the caller's side). */
int apv_device_ptr(apv_handle* h, int tensor_id, void** ptr);
int apv_synchronize(apv_handle* h);

/* Per-stage device time of the last apv_process_block* call, milliseconds:
 * [0] S1 rir_conv [1] S2+S3 wola_weight [2] S4 stats [3] S5 jdiag [4] S6 sweep [5] S7 render [6] total. */
int apv_stage_times(apv_handle* h, float* ms7);
/* Device time of the phases of the last joint diagonalisation (S5), milliseconds:
 * [0] Cholesky [1] two-sided reduction C = L^-1 R_B L^-T [2] tridiagonalisation
 * [3] bisection + inverse iteration [4] back-transformation [5] U = L^-T Q. */
int apv_jdiag_phase_times(apv_handle* h, float* ms6);
/* Device time of the dominant kernels in the last block, milliseconds: [1] syrk_toeplitz_kernel (statistics, FP64
 * tensor bound).  One-stage tridiagonalisation (eig_mode 1): [0] sum over all td_panel_kernel launches (HBM/L2
 * bound), [2] their number.  Two-stage (eig_mode 3 / auto for n >= 1024): [0] dense -> band (DMMA + cluster QR),
 * [3] band -> tridiagonal (bulge chasing), [2] = -1. */
int apv_kernel_times(apv_handle* h, float* ms4);
/* CUDA-event timer on the handle's stream (the stream every kernel of the handle is launched on). */
int apv_timer_start(apv_handle* h);
int apv_timer_stop(apv_handle* h, float* ms);
/* Number of kernel launches issued by the last apv_process_block* call. */
int apv_launch_count(const apv_handle* h);

/* Stand-alone joint diagonalisation = reference jdiag(A, B) (apvast.py:20-36), top-V pairs.
 * A, B: (n, n) host; lambda_out: (V); U_out: (V, n) row v = U[:, v].  pivot_out: first bad pivot if ENOTPD. */
int apv_jdiag(int n, int V, const double* A, const double* B, double reg, int eig_mode, double* lambda_out,
              double* U_out, int* pivot_out);

/* Test / measurement utilities. */
int apv_util_gemm(int M, int N, int K, int transA, int transB, double alpha, const double* A, const double* B,
                  double beta, double* C);
int apv_util_fft(int n, int inverse, const double* in_ri, double* out_ri);
/* FP64 DMMA issue-rate microbenchmark: returns achieved TFLOP/s. */
int apv_bench_dmma_peak(int iters, double* tflops);
/* CUDA-core FP64 FMA: out3 = {chip TFLOP/s, cycles per dependent DFMA, warp issue interval per sub-partition}. */
int apv_bench_dfma(int iters, double* out3);
/* Times nrep launches of the internal GEMM (M=N=K=n) on device data; returns ms per launch. */
int apv_bench_gemm(int n, int nrep, float* ms);
/* One shape of the same building block: C (M x N) = alpha A B(^T) + beta C with K columns, `batch` independent
 * problems, optionally only the lower tiles (tri) stored to both triangles (mirror); bn = forced tile width (0 = auto).
 * Shapes of the joint diagonalisation: rank-64 band updates, rank-256 trailing updates, skinny products. */
int apv_bench_gemm_shape(int M, int N, int K, int batch, int transB, int tri, int mirror, int bn, double beta,
                         int nrep, float* ms);

const char* apv_last_error(void);
const char* apv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* APVAST_B200_H */
