"""Synthetic workloads of the named shapes (BASELINE.json configs).

RIRs: exponentially decaying Gaussian noise, T60 = 80 ms, scale 1e-3 (the magnitude of the reference's
``Python/rirs.mat``); programme signals: "music + speech"-like = sum of AR(2) resonances plus
amplitude-modulated pink-ish noise, unit RMS.  Deterministic (seeded)."""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (L, M, J, K, Nb, N, V, d, seconds)
    "cfg2": dict(L=8, M=8, J=128, K=1024, Nb=2048, N=2048, V=64, d=32, seconds=10.0),
    "cfg3": dict(L=16, M=16, J=256, K=1024, Nb=2048, N=2048, V=64, d=32, seconds=60.0),
    "small": dict(L=4, M=4, J=32, K=128, Nb=256, N=384, V=16, d=8, seconds=0.2),
    # BASELINE configs[4] array size (L=32, J=256, n=8192) as a 2-zone problem (the reference has no 4-zone form)
    "cfg5_2zone": dict(L=32, M=16, J=256, K=1024, Nb=2048, N=2048, V=64, d=32, seconds=1.0),
}


def _rirs(seed, K, L, M, fs=48000.0, t60=0.08):
    rng = np.random.default_rng(seed)
    t = np.arange(K) / fs
    return 1e-3 * rng.standard_normal((K, L, M)) * np.exp(-6.91 * t / t60).reshape(-1, 1, 1)


def _programme(seed, n, fs=48000.0):
    """AR(2) resonances + amplitude-modulated low-passed noise, unit RMS."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    x = np.zeros(n)
    for f0, r in ((220.0, 0.999), (660.0, 0.998), (1500.0, 0.995), (3200.0, 0.99)):
        w0 = 2 * np.pi * f0 / fs
        x += lfilter([1.0], [1.0, -2 * r * np.cos(w0), r * r], rng.standard_normal(n)) * (1 - r)
    noise = lfilter([1.0], [1.0, -0.95], rng.standard_normal(n))
    env = 0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * np.arange(n) / fs + rng.uniform(0, 6.28))
    x = x / np.std(x) + 0.5 * env * noise / np.std(noise)
    return x / np.sqrt(np.mean(x * x))


def make_workload(name: str, n_blocks: int = None, variant: int = 0):
    """`variant` > 0 selects another deterministic draw of the RIRs and the programme signals (parity fixtures)."""
    c = CONFIGS[name]
    vs = 100 * int(variant)
    H = c["Nb"] // 2
    total = int(c["seconds"] * 48000) // H if n_blocks is None else n_blocks
    n = total * H
    cfg = dict(block_size=c["Nb"], filter_length=c["J"], modeling_delay=c["d"], reference_index_A=0,
               reference_index_B=0, number_of_eigenvectors=c["V"], mu=1.0, statistics_buffer_length=c["N"])
    return dict(name=name, cfg=cfg, rir_A=_rirs(10 + vs, c["K"], c["L"], c["M"]),
                rir_B=_rirs(11 + vs, c["K"], c["L"], c["M"]),
                signal_A=_programme(1 + vs, n), signal_B=_programme(2 + vs, n), n_blocks=total, hop=H,
                shapes=dict(c, H=H, n=c["L"] * c["J"]))
