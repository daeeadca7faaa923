"""I/O around the block engine: RIR files, programme-signal feeders, whole-signal
rendering.  Host-side helpers only -- none of this is on the per-block path.

Reference pointers: ``load("rirs.mat")`` / ``load("signals.mat")`` and the per-hop driver loop of
``Python/make_python_test.m:4,19-24,44-54``; ``audioread`` / ``audiowrite`` of ``Matlab/main.m:23-26,35``.
"""
from __future__ import annotations

import numpy as np


def load_rirs_mat(path: str, key_A: str = "rirA", key_B: str = "rirB"):
    """RIRs of the two zones from a MATLAB v5 ``.mat`` file in the reference's layout (``Python/rirs.mat``:
    ``rirA`` and ``rirB``, each (K, L, M) = (taps, loudspeakers, microphones), make_python_test.m:4-6).
    Returns C-contiguous float64 arrays."""
    from scipy.io import loadmat

    d = loadmat(path)
    if key_A not in d or key_B not in d:
        raise KeyError(f"{path}: expected variables {key_A!r} and {key_B!r}, found "
                       f"{sorted(k for k in d if not k.startswith('__'))}")
    rA = np.ascontiguousarray(d[key_A], dtype=np.float64)
    rB = np.ascontiguousarray(d[key_B], dtype=np.float64)
    if rA.ndim != 3 or rA.shape != rB.shape:
        raise RuntimeError("rirs of unequal size")          # the constructor's own check (apvast.py:89-90)
    return rA, rB


def read_wav_mono(path: str):
    """(sampling_rate, float64 samples in [-1, 1)) of a PCM / float WAV file; multi-channel files are averaged."""
    from scipy.io import wavfile

    fs, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float64) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":                                 # 8-bit PCM is unsigned
        x = (x.astype(np.float64) - 128.0) / 128.0
    else:
        x = x.astype(np.float64)
    if x.ndim == 2:
        x = x.mean(axis=1)
    return int(fs), x


def hop_blocks(signal_A, signal_B, hop: int, pad_last: bool = True):
    """Generator of (input_A, input_B) hops for ``process_input_buffers`` from two programme signals (the reshape
    of make_python_test.m:33-38 as a stream).  The shorter signal is zero-padded; with ``pad_last`` the final
    partial hop is zero-padded, otherwise dropped."""
    a = np.asarray(signal_A, dtype=np.float64).reshape(-1)
    b = np.asarray(signal_B, dtype=np.float64).reshape(-1)
    n = max(a.size, b.size)
    nblk = (n + hop - 1) // hop if pad_last else n // hop
    for t in range(nblk):
        out = []
        for x in (a, b):
            seg = x[t * hop:(t + 1) * hop]
            if seg.size < hop:
                seg = np.concatenate([seg, np.zeros(hop - seg.size)])
            out.append(np.ascontiguousarray(seg))
        yield out[0], out[1]


def render_signal(engine, signal_A, signal_B, rank: int = None, collect_filters: bool = False):
    """Run a whole pair of programme signals through ``engine`` hop by hop (the driver loop of
    make_python_test.m:44-54 / main.m:52-62) and return the loudspeaker feeds of one rank.

    rank: 1-based number of eigenvectors of the returned solution (default: the engine's largest).
    Returns (feeds_A, feeds_B) of shape (n_blocks * hop, L) -- ``None`` for a zone that is switched off -- and, with
    ``collect_filters``, the per-block filters (n_blocks, n) of that rank for both zones."""
    V = engine.number_of_eigenvectors
    v = (V if rank is None else int(rank)) - 1
    if not 0 <= v < V:
        raise ValueError(f"rank must be in 1..{V}")
    outA, outB, wA, wB = [], [], [], []
    for a, b in hop_blocks(signal_A, signal_B, engine.hop_size):
        oA, oB, _, _ = engine.process_input_buffers(a, b)
        if oA is not None:
            outA.append(np.array(oA[v]))
        if oB is not None:
            outB.append(np.array(oB[v]))
        if collect_filters:
            if engine.w_A is not None:
                wA.append(np.array(engine.w_A[v]).reshape(-1))
            if engine.w_B is not None:
                wB.append(np.array(engine.w_B[v]).reshape(-1))
    fa = np.concatenate(outA, axis=0) if outA else None
    fb = np.concatenate(outB, axis=0) if outB else None
    if collect_filters:
        return fa, fb, (np.array(wA) if wA else None), (np.array(wB) if wB else None)
    return fa, fb


def write_wav(path: str, sampling_rate: int, feeds, peak: float = None):
    """Loudspeaker feeds (n, L) -> 32-bit float WAV with L channels; ``peak`` rescales to that absolute maximum."""
    from scipy.io import wavfile

    x = np.asarray(feeds, dtype=np.float64)
    if peak is not None:
        m = np.max(np.abs(x))
        if m > 0:
            x = x * (peak / m)
    wavfile.write(path, int(sampling_rate), x.astype(np.float32))
