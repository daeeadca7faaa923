"""Evaluation metrics of the callers' side of the path: predicted pressures
(``Matlab/ControlMethods/predictPressure.m:12-17``), acoustic contrast and normalised mean-square error /
normalised signal distortion (``Matlab/main.m:120-130``).  Host NumPy: these are evaluation utilities, not part
of the per-block hot path."""
from __future__ import annotations

import numpy as np
from scipy.signal import lfilter


def predict_pressure(loudspeaker_signals, rirs):
    """sum_l filter(rirs[:, l, m], 1, feeds[:, l]) -> (T, M)   (predictPressure.m:12-17)."""
    x = np.asarray(loudspeaker_signals, dtype=np.float64)
    T, L = x.shape
    K, L2, M = rirs.shape
    assert L == L2
    p = np.zeros((T, M))
    for m in range(M):
        for l in range(L):
            p[:, m] += lfilter(rirs[:, l, m], 1.0, x[:, l])
    return p


def acoustic_contrast_db(p_bright, p_dark):
    """10 log10(||p_bright||_F^2 / ||p_dark||_F^2)   (main.m:129-130)."""
    return 10.0 * np.log10(np.sum(p_bright ** 2) / np.sum(p_dark ** 2))


def nmse(p_target, p):
    """mean over microphones of ||target - p||^2 / ||target||^2   (main.m:120-127)."""
    num = np.sum((p_target - p) ** 2, axis=0)
    den = np.sum(p_target ** 2, axis=0)
    return float(np.mean(num / den))


def normalised_signal_distortion_db(p_target, p):
    return 10.0 * np.log10(nmse(p_target, p))


def evaluate_zone(feeds, rir_bright, rir_dark, target_signal, reference_index, modeling_delay):
    """AC and NSD of one zone's loudspeaker feeds (T, L).  The target pressure is the programme signal through the
    reference loudspeaker's RIR delayed by the modelling delay (apvast.py:102-112)."""
    K = rir_bright.shape[0]
    tr = np.zeros((K, rir_bright.shape[2]))
    tr[modeling_delay:, :] = rir_bright[:K - modeling_delay, reference_index, :]
    T = feeds.shape[0]
    p_t = np.stack([lfilter(tr[:, m], 1.0, target_signal[:T]) for m in range(tr.shape[1])], axis=1)
    p_b = predict_pressure(feeds, rir_bright)
    p_d = predict_pressure(feeds, rir_dark)
    return acoustic_contrast_db(p_b, p_d), normalised_signal_distortion_db(p_t, p_b)
