"""Drop-in ``apvast`` class and ``jdiag`` function backed by the sm_100a CUDA engine.

Mirrors the public surface of the reference ``Python/apvast.py``:

* constructor ``apvast(block_size, rir_A, rir_B, filter_length, modeling_delay, reference_index_A,
  reference_index_B, number_of_eigenvectors, mu, statistics_buffer_length, hop_size=None,
  sampling_rate=48000, run_A=True, run_B=True, perceptual=True)``           (reference :40-56)
* per-block call ``process_input_buffers(input_A, input_B) -> (out_A, out_B, out_A_t, out_B_t)``,
  each a list of V arrays (H, L), ``None`` for a zone that is switched off      (reference :153-165)
* observable attributes after a call: ``w_A, w_B`` (V, n, 1), ``lambda_A/B``, ``U_A/B``,
  ``R_A_to_A ... R_B_to_B`` (n, n), ``r_A, r_B`` (n, 1), ``weighting_spectra_A/B`` (F, M) complex,
  the state buffers of reference :115-151 -- fetched lazily from the device on attribute access.
* errors: ``RuntimeError("block size must be modulo 2")``, ``RuntimeError("rirs of unequal size")``,
  ``RuntimeError("invalid input size")``, ``numpy.linalg.LinAlgError`` when R_D + reg I is not
  positive definite                                                     (reference :86-90,154-155,21-24)

Differences that are supersets of the reference contract:
  - returned output arrays are fresh host arrays, not views into overlap buffers;
  - only the V leading joint eigenpairs are formed (all the filters need): ``U_X`` is (n, V),
    ``lambda_X`` is (V,);
  - ``input_B`` is size-checked too;
  - extra keyword-only arguments (``model``, ``device``, ``eig_mode``) and extra methods
    (``sweep``, ``process_signal``, ``get_state``/``set_state``, ``stage_times``).

All arithmetic runs in hand-written CUDA kernels behind the C-ABI of ``include/apvast_b200.h``;
there is no NumPy/CPU fallback for any stage of the per-block path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi

# module-level switches with the reference's names and defaults (Python/apvast.py:6-7);
# read when an engine is constructed
EXPERIMENTAL_NORMALIZE_GAINS = True
EXPERIMENTAL_REGULARIZATION = True

_PATHS = ("A_to_A", "A_to_B", "B_to_A", "B_to_B")


def jdiag(A, B, number_of_eigenvectors=None, reg=1e-7, eig_mode=0):
    """Joint diagonalisation on the GPU: reference ``jdiag(A, B)`` (``Python/apvast.py:20-36``).

    Returns ``(U, D)``: ``U`` (n, V) with ``U.T (B + reg I) U = I`` and ``U.T A U = D``, ``D`` dense
    diagonal (V, V), eigenvalues descending.  ``V`` defaults to n (the reference always forms all n).
    Raises ``numpy.linalg.LinAlgError`` when ``B + reg I`` is not positive definite.
    """
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    n = A.shape[0]
    if A.shape != (n, n) or B.shape != (n, n):
        raise RuntimeError("jdiag expects two square matrices of equal size")
    V = n if number_of_eigenvectors is None else int(number_of_eigenvectors)
    lam = np.zeros(V)
    Ut = np.zeros((V, n))
    piv = C.c_int(0)
    capi.check(capi.lib().apv_jdiag(n, V, capi.ptr(A), capi.ptr(B), float(reg), int(eig_mode), capi.ptr(lam),
                                    capi.ptr(Ut), C.byref(piv)))
    return np.ascontiguousarray(Ut.T), np.diag(lam)


class apvast:
    def __init__(self, block_size: int, rir_A, rir_B, filter_length: int, modeling_delay: int,
                 reference_index_A: int, reference_index_B: int, number_of_eigenvectors: int, mu: float,
                 statistics_buffer_length: int, hop_size: int = None, sampling_rate: int = 48000,
                 run_A: bool = True, run_B: bool = True, perceptual: bool = True, *, model=None,
                 device: int = None, eig_mode: int = 0, stats_mode: int = 0, flavour: str = "python",
                 fullscale_db: float = 94.0, active_mics_A: int = 0, split_weights: bool = False):
        self._h = None
        if flavour not in ("python", "matlab"):
            raise RuntimeError("flavour must be 'python' or 'matlab'")
        self.flavour = flavour
        self.block_size = int(block_size)
        self.rir_A = rir_A
        self.rir_B = rir_B
        self.filter_length = int(filter_length)
        self.modeling_delay = int(modeling_delay)
        self.reference_index_A = int(reference_index_A)
        self.reference_index_B = int(reference_index_B)
        # the MATLAB class takes an ascending LIST of ranks and returns one solution per element (apVast.m:204,527-544);
        # the Python class takes an int V and returns ranks 1..V.  Both are accepted.
        self._ranks = None
        if np.ndim(number_of_eigenvectors) > 0:
            self._ranks = [int(r) for r in np.ravel(number_of_eigenvectors)]
            if self._ranks != sorted(self._ranks) or self._ranks[0] < 1:
                raise RuntimeError("the list of ranks must be ascending and >= 1")
            number_of_eigenvectors = self._ranks[-1]
        self.number_of_eigenvectors = int(number_of_eigenvectors)
        self._mu = float(mu)
        self.sampling_rate = sampling_rate
        self.statistics_buffer_length = int(statistics_buffer_length)
        self.run_A = bool(run_A)
        self.run_B = bool(run_B)
        self.perceptual = bool(perceptual)

        # validate exactly like the reference (:86-90)
        if self.block_size % 2 != 0:
            raise RuntimeError("block size must be modulo 2")
        if np.shape(rir_A) != np.shape(rir_B):
            raise RuntimeError("rirs of unequal size")

        rA = np.ascontiguousarray(rir_A, dtype=np.float64)
        rB = np.ascontiguousarray(rir_B, dtype=np.float64)
        if rA.ndim != 3:
            raise RuntimeError("rirs must have shape (rir_length, number_of_srcs, number_of_mics)")
        self.hop_size = int(hop_size) if hop_size else self.block_size // 2
        self.window = np.sin(np.pi / self.block_size * np.arange(self.block_size)).reshape(-1, 1)
        self.rir_length, self.number_of_srcs, self.number_of_mics = rA.shape
        Nb, K, L, M = self.block_size, self.rir_length, self.number_of_srcs, self.number_of_mics

        # host perceptual model: device tables (default) or a host callable with .gain() (e.g. libdetectability)
        self.model = None
        mode = 0
        if self.perceptual:
            if model is None:
                from .perceptual import MaskingModel
                self.model = MaskingModel(Nb, sampling_rate, fullscale_db)
                mode = 3 if split_weights else 1      # 3: S1 + S2, then weighting curves may be exchanged (zones.py)
            else:
                self.model = model
                mode = 2
            np.seterr(divide="ignore")       # as the reference does (:76)

        # the six randn start buffers, drawn from the global NumPy RNG in the reference order (:124-129)
        # (the MATLAB class starts from zeros, apVast.m:175-180, and does not touch the RNG)
        init = None
        if flavour == "python":
            init = np.concatenate([(1e-3 * np.random.randn(Nb, L, M)).ravel() for _ in range(4)] +
                                  [(1e-3 * np.random.randn(Nb, M)).ravel() for _ in range(2)])
        matlab = flavour == "matlab"

        cfg = capi.Config(
            block_size=Nb, hop_size=self.hop_size, rir_length=K, n_srcs=L, n_mics=M,
            filter_length=self.filter_length, stats_length=self.statistics_buffer_length,
            n_eig=self.number_of_eigenvectors, modeling_delay=self.modeling_delay,
            ref_A=self.reference_index_A, ref_B=self.reference_index_B, run_A=int(self.run_A), run_B=int(self.run_B),
            perceptual=mode, normalize_gains=2 if matlab else int(bool(EXPERIMENTAL_NORMALIZE_GAINS)),
            eig_mode=int(eig_mode),
            stats_mode=int(stats_mode), device=-1 if device is None else int(device), mu=self._mu, reg=1e-7,
            sampling_rate=float(sampling_rate), toeplitz_clean=int(matlab), normalize_stats=int(matlab),
            loading_mode=int(matlab), target_ref_per_zone=int(matlab), bright_load=1e-8, dark_load=5e-3,
            active_mics_A=int(active_mics_A), reg_relative=int(not EXPERIMENTAL_REGULARIZATION))
        h = C.c_void_p()
        capi.check(capi.lib().apv_create(C.byref(cfg), capi.ptr(rA), capi.ptr(rB), capi.ptr(init), C.byref(h)))
        self._h = h
        if mode in (1, 3):
            capi.check(capi.lib().apv_set_gain_table(self._h, self.model.n_channels, capi.ptr(self.model.G2),
                                                     self.model.Cs, self.model.Ca, self.model.Leff))
        self._mode = mode
        self._n = self.filter_length * L
        self._blocks = 0
        self._reg_relative = not EXPERIMENTAL_REGULARIZATION
        self._between = None       # mode 3: hook called between S2 and S3 (exchange of weighting curves)

    def _sync_flags(self):
        # the reference reads EXPERIMENTAL_REGULARIZATION inside jdiag, i.e. at call time (Python/apvast.py:22-27)
        rel = not EXPERIMENTAL_REGULARIZATION
        if rel != self._reg_relative:
            capi.check(capi.lib().apv_set_reg_mode(self._h, int(rel)))
            self._reg_relative = rel

    # ------------------------------------------------------------------ life-cycle
    def close(self):
        if getattr(self, "_h", None):
            capi.lib().apv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ mu is read at call time (:161)
    @property
    def mu(self):
        return self._mu

    @mu.setter
    def mu(self, value):
        self._mu = float(value)
        if self._h:
            capi.check(capi.lib().apv_set_mu(self._h, self._mu))

    # ------------------------------------------------------------------ per-block call (:153-165)
    def process_input_buffers(self, input_A, input_B):
        if np.size(input_A) != self.hop_size or np.size(input_B) != self.hop_size:
            raise RuntimeError("invalid input size")
        a = np.ascontiguousarray(input_A, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(input_B, dtype=np.float64).reshape(-1)
        V, H, L = self.number_of_eigenvectors, self.hop_size, self.number_of_srcs
        oA = np.empty((V, H, L)) if self.run_A else None
        oB = np.empty((V, H, L)) if self.run_B else None
        oAt = np.empty((H, L))
        oBt = np.empty((H, L))
        lib = capi.lib()
        self._sync_flags()
        if self._mode in (2, 3):
            capi.check(lib.apv_begin_block(self._h, capi.ptr(a), capi.ptr(b)))
            if self._mode == 2:
                self._host_gains()
            elif self._between is not None:
                self._between(self)
            capi.check(lib.apv_finish_block(self._h, capi.ptr(oA), capi.ptr(oB), capi.ptr(oAt), capi.ptr(oBt)))
        else:
            capi.check(lib.apv_process_block(self._h, capi.ptr(a), capi.ptr(b), capi.ptr(oA), capi.ptr(oB),
                                             capi.ptr(oAt), capi.ptr(oBt)))
        self._blocks += 1
        sel = range(V) if self._ranks is None else [r - 1 for r in self._ranks]
        out_A = [oA[v] for v in sel] if self.run_A else None
        out_B = [oB[v] for v in sel] if self.run_B else None
        # the reference returns V identical target arrays (:418,422,467-475,501,504)
        return out_A, out_B, [oAt for _ in sel], [oBt for _ in sel]

    def _begin(self, input_A, input_B):
        """S1 + S2 of a split block (mode 3); ``_finish`` runs S3..S7."""
        if np.size(input_A) != self.hop_size or np.size(input_B) != self.hop_size:
            raise RuntimeError("invalid input size")
        a = np.ascontiguousarray(input_A, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(input_B, dtype=np.float64).reshape(-1)
        self._sync_flags()
        capi.check(capi.lib().apv_begin_block(self._h, capi.ptr(a), capi.ptr(b)))

    def _finish(self):
        V, H, L = self.number_of_eigenvectors, self.hop_size, self.number_of_srcs
        oA = np.empty((V, H, L)) if self.run_A else None
        oB = np.empty((V, H, L)) if self.run_B else None
        oAt, oBt = np.empty((H, L)), np.empty((H, L))
        capi.check(capi.lib().apv_finish_block(self._h, capi.ptr(oA), capi.ptr(oB), capi.ptr(oAt), capi.ptr(oBt)))
        self._blocks += 1
        sel = range(V) if self._ranks is None else [r - 1 for r in self._ranks]
        return ([oA[v] for v in sel] if self.run_A else None, [oB[v] for v in sel] if self.run_B else None,
                [oAt for _ in sel], [oBt for _ in sel])

    def _host_gains(self):
        """perceptual with a host model: W[:, m] = model.gain(time block), unit-norm (:313-324)."""
        M, Nb, F = self.number_of_mics, self.block_size, self.block_size // 2 + 1
        frames = self._get(capi.T_TARGET_FRAME).reshape(2, M, Nb)
        W = np.empty((2, M, F))
        for z in range(2):
            for m in range(M):
                g = np.real(np.asarray(self.model.gain(frames[z, m]), dtype=complex))
                if self.flavour == "matlab":
                    g = g / np.sqrt(np.sum(g * g) + np.sum(g[1:-1] ** 2))
                elif EXPERIMENTAL_NORMALIZE_GAINS:
                    g = g / np.linalg.norm(g)
                W[z, m] = g
        self._set(capi.T_WEIGHT, W)

    def process_blocks(self, signal_A, signal_B, want_filters=False):
        """Throughput form of the per-hop loop (``make_python_test.m:44-54``): ``nblocks`` consecutive hops in ONE call,
        results identical to ``nblocks`` calls of ``process_input_buffers``.  S1-S4 of hop t+1 overlap S5-S7 of hop t
        on the device and the rendered hops are copied out asynchronously.

        Returns ``(out_A, out_B, out_A_t, out_B_t[, w])``: arrays (nblocks, V, H, L) (``None`` for a zone that is off),
        (nblocks, H, L) for the target streams, and with ``want_filters`` the per-hop filters (nblocks, 2, V, n)."""
        if self._mode in (2, 3):
            raise RuntimeError("process_blocks is not available with a split perceptual call")
        a = np.ascontiguousarray(signal_A, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(signal_B, dtype=np.float64).reshape(-1)
        H = self.hop_size
        if a.size != b.size or a.size % H != 0:
            raise RuntimeError("invalid input size")
        nb = a.size // H
        V, L = self.number_of_eigenvectors, self.number_of_srcs
        oA = np.empty((nb, V, H, L)) if self.run_A else None
        oB = np.empty((nb, V, H, L)) if self.run_B else None
        oAt, oBt = np.empty((nb, H, L)), np.empty((nb, H, L))
        w = np.empty((nb, 2, V, self._n)) if want_filters else None
        self._sync_flags()
        capi.check(capi.lib().apv_process_blocks(self._h, nb, capi.ptr(a), capi.ptr(b), capi.ptr(oA), capi.ptr(oB),
                                                 capi.ptr(oAt), capi.ptr(oBt), capi.ptr(w)))
        self._blocks += nb
        if self._ranks is not None:
            sel = [r - 1 for r in self._ranks]
            oA = oA[:, sel] if oA is not None else None
            oB = oB[:, sel] if oB is not None else None
        return (oA, oB, oAt, oBt, w) if want_filters else (oA, oB, oAt, oBt)

    def set_pipeline(self, mode):
        """Overlap of S1-S4 of hop t+1 with S5-S7 of hop t in multi-hop calls: 0 / False off, 1 from the start of S5,
        2 / True (default) from the bulge chasing of hop t on.  Results are identical."""
        capi.check(capi.lib().apv_set_pipeline(self._h, 2 if mode is True else int(mode)))

    def set_depth(self, depth: int):
        """Joint diagonalisations in flight in multi-hop calls (1 or 2; default 2 for n < 2048).  Results are identical."""
        capi.check(capi.lib().apv_set_depth(self._h, int(depth)))

    def advance_state(self, input_A, input_B):
        """S1-S3 only (state update without statistics/filters/rendering): warm-up of a block range."""
        if self._mode in (2, 3):
            raise RuntimeError("advance_state is not available with a split perceptual call")
        a = np.ascontiguousarray(input_A, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(input_B, dtype=np.float64).reshape(-1)
        if a.size != self.hop_size or b.size != self.hop_size:
            raise RuntimeError("invalid input size")
        capi.check(capi.lib().apv_advance_state(self._h, capi.ptr(a), capi.ptr(b)))

    def sweep(self, mu_values):
        """mu x V trade-off sweep from the joint diagonalisation of the last block (BASELINE cfg-4).

        Returns (w_A, w_B), each (n_mu, V, n) (``None`` for a zone that is off)."""
        mus = np.ascontiguousarray(mu_values, dtype=np.float64).reshape(-1)
        out = np.zeros((mus.size, 2, self.number_of_eigenvectors, self._n))
        capi.check(capi.lib().apv_sweep(self._h, mus.size, capi.ptr(mus), capi.ptr(out)))
        return (out[:, 0] if self.run_A else None), (out[:, 1] if self.run_B else None)

    def sweep_metrics(self, mu_values, device_out=None):
        """Eigen-basis figures of merit of the mu x V sweep, (n_mu, 2, V, 3): dark energy ``w'(R_D + reg I)w``, bright
        energy ``w'R_B w`` and ``w'r_B`` of every rank's filter, from the joint diagonalisation of the last block.
        ``device_out``: optional device pointer of an (n_mu, 2, V, n) buffer that receives the filters themselves."""
        mus = np.ascontiguousarray(mu_values, dtype=np.float64).reshape(-1)
        out = np.zeros((mus.size, 2, self.number_of_eigenvectors, 3))
        capi.check(capi.lib().apv_sweep_device(self._h, mus.size, capi.ptr(mus),
                                               None if device_out is None else C.c_void_p(device_out), capi.ptr(out)))
        return out

    def evaluate(self, feeds, signal, zone="A"):
        """Acoustic contrast and normalised signal distortion of loudspeaker feeds (T, L) at the control microphones,
        on the device (``predictPressure.m:12-17``, ``main.m:120-130``).  Returns (AC dB, NMSE, NSD dB)."""
        f = np.ascontiguousarray(feeds, dtype=np.float64)
        s = np.ascontiguousarray(signal, dtype=np.float64).reshape(-1)[:f.shape[0]]
        if f.ndim != 2 or f.shape[1] != self.number_of_srcs or s.size != f.shape[0]:
            raise RuntimeError("feeds must be (T, number_of_srcs) and signal at least T long")
        out = np.zeros(3)
        capi.check(capi.lib().apv_eval_zone(self._h, 0 if zone in ("A", 0) else 1, f.shape[0], capi.ptr(f), capi.ptr(s),
                                            capi.ptr(out)))
        return float(out[0]), float(out[1]), float(out[2])

    def stage_times(self):
        """Device milliseconds of the last block: dict S1, S2S3, S4, S5, S6, S7, total; and launch count."""
        ms = (C.c_float * 7)()
        capi.check(capi.lib().apv_stage_times(self._h, ms))
        keys = ("S1_rir_conv", "S2S3_wola_weight", "S4_stats", "S5_jdiag", "S6_sweep", "S7_render", "total")
        d = {k: float(ms[i]) for i, k in enumerate(keys)}
        d["launches"] = int(capi.lib().apv_launch_count(self._h))
        ph = (C.c_float * 6)()
        capi.check(capi.lib().apv_jdiag_phase_times(self._h, ph))
        for i, k in enumerate(("chol", "reduce", "tridiag", "eig", "backtransform", "backsolve")):
            d["S5_" + k] = float(ph[i])
        return d

    # ------------------------------------------------------------------ raw tensor access
    def _get(self, tid):
        n = capi.lib().apv_tensor_size(self._h, tid)
        a = np.empty(n)
        capi.check(capi.lib().apv_get(self._h, tid, capi.ptr(a), n))
        return a

    def _set(self, tid, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        capi.check(capi.lib().apv_set(self._h, tid, capi.ptr(a), a.size))

    def _sel(self, w):
        return w.copy() if self._ranks is None else w[[r - 1 for r in self._ranks]].copy()

    def _zone(self, z, on):
        if not on or self._blocks == 0:
            return None
        return z

    # ------------------------------------------------------------------ observable results
    @property
    def w_A(self):
        if not self.run_A:
            return None
        V, n = self.number_of_eigenvectors, self._n
        return self._sel(self._get(capi.T_W).reshape(2, V, n)[0].reshape(V, n, 1))

    @property
    def w_B(self):
        if not self.run_B:
            return None
        V, n = self.number_of_eigenvectors, self._n
        return self._sel(self._get(capi.T_W).reshape(2, V, n)[1].reshape(V, n, 1))

    @property
    def lambda_A(self):
        return self._get(capi.T_LAMBDA).reshape(2, -1)[0].copy() if self.run_A else None

    @property
    def lambda_B(self):
        return self._get(capi.T_LAMBDA).reshape(2, -1)[1].copy() if self.run_B else None

    @property
    def U_A(self):
        V, n = self.number_of_eigenvectors, self._n
        return self._get(capi.T_U).reshape(2, V, n)[0].T.copy() if self.run_A else None

    @property
    def U_B(self):
        V, n = self.number_of_eigenvectors, self._n
        return self._get(capi.T_U).reshape(2, V, n)[1].T.copy() if self.run_B else None

    def _R(self, p):
        n = self._n
        return self._get(capi.T_R).reshape(4, n, n)[p].copy()

    R_A_to_A = property(lambda self: self._R(0) if self.run_A else None)
    R_A_to_B = property(lambda self: self._R(1) if self.run_A else None)
    R_B_to_A = property(lambda self: self._R(2) if self.run_B else None)
    R_B_to_B = property(lambda self: self._R(3) if self.run_B else None)

    @property
    def r_A(self):
        return self._get(capi.T_RVEC).reshape(2, -1)[0].reshape(-1, 1).copy() if self.run_A else None

    @property
    def r_B(self):
        return self._get(capi.T_RVEC).reshape(2, -1)[1].reshape(-1, 1).copy() if self.run_B else None

    def _weights(self, z):
        M, F = self.number_of_mics, self.block_size // 2 + 1
        return self._get(capi.T_WEIGHT).reshape(2, M, F)[z].T + 0j      # (F, M) complex like the reference

    weighting_spectra_A = property(lambda self: self._weights(0))
    weighting_spectra_B = property(lambda self: self._weights(1))

    def _filter_spectra(self, w):
        J, L, Nb = self.filter_length, self.number_of_srcs, self.block_size
        return [np.fft.rfft(w[v, :, 0].reshape(L, J).T, Nb, axis=0) for v in range(w.shape[0])]

    # filter spectra are a pure re-encoding of w (reference :417-422); provided for observability only --
    # the engine renders in the time domain and never needs them
    filter_spectra_A = property(lambda self: self._filter_spectra(self.w_A) if self.run_A else None)
    filter_spectra_B = property(lambda self: self._filter_spectra(self.w_B) if self.run_B else None)

    def _target_spectra(self):
        J, L, Nb, V = self.filter_length, self.number_of_srcs, self.block_size, self.number_of_eigenvectors
        t = np.zeros(J * L)
        t[J * self.reference_index_A + self.modeling_delay] = 1.0
        ft = np.fft.rfft(t.reshape(L, J).T, Nb, axis=0)
        return [ft for _ in range(V)]

    filter_spectra_A_t = property(lambda self: self._target_spectra())
    filter_spectra_B_t = property(lambda self: self._target_spectra())

    # ------------------------------------------------------------------ state buffers (reference :115-151)
    def _paths(self, tid, T):
        M, L = self.number_of_mics, self.number_of_srcs
        return self._get(tid).reshape(4, M, L, T)

    def _path_attr(self, tid, T, p):
        return np.ascontiguousarray(self._paths(tid, T)[p].transpose(2, 1, 0))        # (T, L, M)

    def _zone_attr(self, tid, T, z):
        M = self.number_of_mics
        return np.ascontiguousarray(self._get(tid).reshape(2, M, T)[z].T)            # (T, M)

    def __getattr__(self, name):
        # lazily materialised state attributes with the reference's names
        if name.startswith("_"):
            raise AttributeError(name)
        Nb, N = self.block_size, self.statistics_buffer_length
        for p, xy in enumerate(_PATHS):
            if name == f"loudspeaker_response_{xy}_buffer":
                return self._path_attr(capi.T_RESP, Nb, p)
            if name == f"loudspeaker_weighted_response_{xy}_overlap_buffer":
                return self._path_attr(capi.T_OLA, Nb, p)
            if name == f"loudspeaker_weighted_response_{xy}_buffer":
                return self._path_attr(capi.T_STATS, N, p)
        for z, xx in enumerate(("A_to_A", "B_to_B")):
            if name == f"loudspeaker_target_response_{xx}_buffer":
                return self._zone_attr(capi.T_RESP_T, Nb, z)
            if name == f"loudspeaker_weighted_target_response_{xx}_overlap_buffer":
                return self._zone_attr(capi.T_OLA_T, Nb, z)
            if name == f"loudspeaker_weighted_target_response_{xx}_buffer":
                return self._zone_attr(capi.T_STATS_T, N, z)
        V, L = self.number_of_eigenvectors, self.number_of_srcs
        if name in ("output_A_overlap_buffer", "output_B_overlap_buffer"):
            z = 0 if name == "output_A_overlap_buffer" else 1
            return np.ascontiguousarray(self._get(capi.T_OUT_OLA).reshape(2, V, L, Nb)[z].transpose(0, 2, 1))
        if name in ("output_A_t_overlap_buffer", "output_B_t_overlap_buffer"):
            z = 0 if name == "output_A_t_overlap_buffer" else 1
            g = self._get(capi.T_OUT_OLA_T).reshape(2, Nb)[z]
            out = np.zeros((V, Nb, L))
            ref = self.reference_index_B if (z == 1 and self.flavour == "matlab") else self.reference_index_A
            lt = (self.filter_length * ref + self.modeling_delay) // self.filter_length
            out[:, :, lt] = g[None, :]
            return out
        if name in ("input_A_block", "input_B_block"):
            x = self._get(capi.T_INPUT).reshape(2, -1)[0 if name == "input_A_block" else 1]
            return x[-Nb:].reshape(-1, 1).copy()
        raise AttributeError(name)

    # ------------------------------------------------------------------ checkpoint / warm start
    _STATE_IDS = (capi.T_RESP, capi.T_RESP_T, capi.T_OLA, capi.T_OLA_T, capi.T_STATS, capi.T_STATS_T,
                  capi.T_OUT_OLA, capi.T_OUT_OLA_T, capi.T_INPUT, capi.T_WEIGHT)

    def get_state(self):
        """Complete streaming state (device layouts) -- what a checkpoint or a warm start needs."""
        return {int(t): self._get(t) for t in self._STATE_IDS}

    def set_state(self, state):
        for t, a in state.items():
            self._set(int(t), a)
        self._blocks = max(self._blocks, 1)

    def save_state(self, path):
        """Checkpoint the streaming state to an .npz file (the reference's closest facility is the property dump
        of Python/make_python_test.m:19-24,55-64)."""
        np.savez(path, **{f"t{t}": a for t, a in self.get_state().items()})

    def load_state(self, path):
        z = np.load(path)
        self.set_state({int(k[1:]): z[k] for k in z.files})
