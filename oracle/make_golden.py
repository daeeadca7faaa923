"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container (where ``/root/reference`` is mounted):

    python -m oracle.make_golden

Each fixture holds the inputs (RIRs, per-hop input samples, NumPy legacy seed used before
construction) and what the reference class ``apvast`` (``Python/apvast.py:39``) produced: filters
``w_A/w_B`` per block, top eigenvalues, statistics, outputs and -- for the small cases -- the complete
state after the last block.  The GPU box has no ``/root/reference``; the ``-m gpu`` parity tests read
only these files.

Cases
  tiny        K=48 L=3 M=2 Nb=64 H=32 J=8 N=96 V=6      (default hop)
  tiny_hop    same, hop_size=16
  tiny_runA   same, run_B=False
  tiny_full   same, V=n=24 (full rank; closed form w = (R_B + mu (R_D + reg I))^-1 r_B applies)
  tiny_perc   same, perceptual=True with the gain model injected at the libdetectability boundary
              (oracle/perceptual_oracle.py; pins everything except the gain formula itself)
  mid         K=96 L=4 M=3 Nb=256 H=128 J=32 N=320 V=16  coloured inputs
  cfg1        Python/rirs.mat, make_python_test.m:6-15 parameters, 10 hops
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference, reference_rirs  # noqa: E402
from oracle.perceptual_oracle import PerceptualModelOracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

STATE_ATTRS = [
    "loudspeaker_response_A_to_A_buffer", "loudspeaker_response_A_to_B_buffer",
    "loudspeaker_response_B_to_A_buffer", "loudspeaker_response_B_to_B_buffer",
    "loudspeaker_target_response_A_to_A_buffer", "loudspeaker_target_response_B_to_B_buffer",
    "loudspeaker_weighted_response_A_to_A_overlap_buffer", "loudspeaker_weighted_response_A_to_B_overlap_buffer",
    "loudspeaker_weighted_response_B_to_A_overlap_buffer", "loudspeaker_weighted_response_B_to_B_overlap_buffer",
    "loudspeaker_weighted_target_response_A_to_A_overlap_buffer",
    "loudspeaker_weighted_target_response_B_to_B_overlap_buffer",
    "loudspeaker_weighted_response_A_to_A_buffer", "loudspeaker_weighted_response_A_to_B_buffer",
    "loudspeaker_weighted_response_B_to_A_buffer", "loudspeaker_weighted_response_B_to_B_buffer",
    "loudspeaker_weighted_target_response_A_to_A_buffer", "loudspeaker_weighted_target_response_B_to_B_buffer",
    "output_A_overlap_buffer", "output_B_overlap_buffer", "output_A_t_overlap_buffer", "output_B_t_overlap_buffer",
    "input_A_block", "input_B_block",
    "R_A_to_A", "R_A_to_B", "R_B_to_A", "R_B_to_B", "r_A", "r_B",
    "weighting_spectra_A", "weighting_spectra_B",
]


def coloured(rng, n, kind):
    x = rng.standard_normal(n)
    if kind == "ar1":
        y = np.zeros(n)
        acc = 0.0
        for i in range(n):
            acc = 0.95 * acc + x[i]
            y[i] = acc
        return 10.0 * y / np.std(y)
    return x


def run_case(name, ref, rirA, rirB, cfg, nblk, seed=0, input_kind="white", ctor_kw=None, full_state=True,
             w_blocks=None, w_ranks=None, out_ranks=None, perceptual_model=False):
    ctor_kw = dict(ctor_kw or {})
    if perceptual_model:
        # inject the gain model where the reference constructs ld.Detectability (apvast.py:77-83)
        class _Det:
            def __init__(self, frame_size, sampling_rate, **kw):
                self._m = PerceptualModelOracle(frame_size, sampling_rate)

            def gain(self, x):
                return self._m.gain(x)

        ref.ld.Detectability = _Det
        ctor_kw["perceptual"] = True
    else:
        ctor_kw["perceptual"] = False
    np.random.seed(seed)
    ap = ref.apvast(rir_A=rirA, rir_B=rirB, **cfg, **ctor_kw)
    H = ap.hop_size
    rng = np.random.default_rng(1)
    sigA = coloured(rng, nblk * H, input_kind)
    sigB = coloured(rng, nblk * H, input_kind)
    V = cfg["number_of_eigenvectors"]
    w_blocks = list(range(nblk)) if w_blocks is None else w_blocks
    w_ranks = list(range(V)) if w_ranks is None else w_ranks
    out_ranks = list(range(V)) if out_ranks is None else out_ranks
    g = {"rir_A": rirA, "rir_B": rirB, "input_A": sigA.reshape(nblk, H), "input_B": sigB.reshape(nblk, H),
         "seed": np.int64(seed), "nblk": np.int64(nblk), "w_blocks": np.array(w_blocks), "w_ranks": np.array(w_ranks),
         "out_ranks": np.array(out_ranks), "hop_size": np.int64(H)}
    for k, v in cfg.items():
        g["cfg_" + k] = np.array(v)
    for k, v in ctor_kw.items():
        g["ctor_" + k] = np.array(v)
    for t in range(nblk):
        outs = ap.process_input_buffers(sigA[t * H:(t + 1) * H], sigB[t * H:(t + 1) * H])
        if t in w_blocks:
            if ap.w_A is not None:
                g[f"w_A_{t}"] = ap.w_A[w_ranks, :, 0].copy()
                g[f"lambda_A_{t}"] = ap.lambda_A[:V].copy()
            if ap.w_B is not None:
                g[f"w_B_{t}"] = ap.w_B[w_ranks, :, 0].copy()
                g[f"lambda_B_{t}"] = ap.lambda_B[:V].copy()
            for nm, o in zip(("out_A", "out_B", "out_A_t", "out_B_t"), outs):
                if o is not None:
                    g[f"{nm}_{t}"] = np.stack([o[v] for v in out_ranks]).copy()
            g[f"r_A_{t}"] = getattr(ap, "r_A", np.zeros(0)).copy() if ap.run_A else np.zeros(0)
            g[f"r_B_{t}"] = getattr(ap, "r_B", np.zeros(0)).copy() if ap.run_B else np.zeros(0)
            # compact statistics pins: diagonal + three rows of every R
            for nm in ("R_A_to_A", "R_A_to_B", "R_B_to_A", "R_B_to_B"):
                R = getattr(ap, nm, None)
                if R is not None:
                    n = R.shape[0]
                    g[f"{nm}_diag_{t}"] = np.diag(R).copy()
                    g[f"{nm}_rows_{t}"] = R[[0, n // 2 - 1, n - 1], :].copy()
    if full_state:
        for a in STATE_ATTRS:
            v = getattr(ap, a, None)
            if v is not None:
                g["state_" + a] = np.array(v)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **g)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    ref = load_reference()
    rng = np.random.default_rng(5)
    K, L, M = 48, 3, 2
    rA = 1e-3 * rng.standard_normal((K, L, M)) * np.exp(-np.arange(K) / 12.0).reshape(-1, 1, 1)
    rB = 1e-3 * rng.standard_normal((K, L, M)) * np.exp(-np.arange(K) / 12.0).reshape(-1, 1, 1)
    tiny = dict(block_size=64, filter_length=8, modeling_delay=3, reference_index_A=1, reference_index_B=2,
                number_of_eigenvectors=6, mu=1.0, statistics_buffer_length=96)
    run_case("tiny", ref, rA, rB, tiny, 8)
    run_case("tiny_hop", ref, rA, rB, tiny, 12, ctor_kw=dict(hop_size=16))
    run_case("tiny_runA", ref, rA, rB, tiny, 6, ctor_kw=dict(run_B=False))
    full = dict(tiny)
    full["number_of_eigenvectors"] = 24
    full["mu"] = 0.3
    run_case("tiny_full", ref, rA, rB, full, 6)
    run_case("tiny_perc", ref, rA, rB, tiny, 8, perceptual_model=True)

    rng = np.random.default_rng(7)
    K, L, M = 96, 4, 3
    dec = np.exp(-np.arange(K) / 20.0).reshape(-1, 1, 1)
    rA = 1e-3 * rng.standard_normal((K, L, M)) * dec
    rB = 1e-3 * rng.standard_normal((K, L, M)) * dec
    mid = dict(block_size=256, filter_length=32, modeling_delay=8, reference_index_A=0, reference_index_B=3,
               number_of_eigenvectors=16, mu=0.5, statistics_buffer_length=320)
    run_case("mid", ref, rA, rB, mid, 8, input_kind="ar1", full_state=False)

    rA, rB = reference_rirs()
    cfg1 = dict(block_size=1600, filter_length=100, modeling_delay=20, reference_index_A=6, reference_index_B=6,
                number_of_eigenvectors=50, mu=1.0, statistics_buffer_length=1000)
    run_case("cfg1", ref, rA, rB, cfg1, 10, full_state=False, w_blocks=[1, 5, 9],
             w_ranks=[0, 1, 2, 5, 10, 20, 30, 40, 49], out_ranks=[0, 49])


if __name__ == "__main__":
    main()
