"""Set-up-time tables of the spectral-integration masking model (van de Par et al. 2005) that the
on-device ``masking_gain`` kernel consumes.

The Python reference takes its gain from the third-party ``libdetectability`` package
(``Python/apvast.py:77-83,318-319``; not vendored, unpinned).  The same model ships in the
reference as MATLAB (``Matlab/ControlMethods/perceptualModel.m:30-139``,
``gammatoneFilterResponse.m``, ``interpolatedThresholdOfHearing.m``); this module builds, once per
engine, what that model needs per block:

    G2[c, f]  = (outer/middle-ear response(f) * gammatone_c(f))**2          (C channels x F bins)
    Cs, Ca    = calibration constants (70 dB SPL masker / 52 dB SPL probe, bisection)
    Leff      = min(Nb / fs / 0.3, 1)

Per block and microphone the kernel then evaluates
    p_c = sum_f G2[c, f] * (2 / Nb**2) |X(f)|**2 ,   gain(f) = sqrt(Cs Leff sum_c G2[c, f] / (p_c + Ca)).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.interpolate import CubicSpline

# ISO 226:2003 threshold in quiet (interpolatedThresholdOfHearing.m:29-30)
ISO226_HZ = (20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800, 1000, 1250,
             1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500)
ISO226_DB = (78.5, 68.7, 59.5, 51.1, 44.0, 37.5, 31.5, 26.5, 22.1, 17.9, 14.4, 11.4, 8.6, 6.2, 4.4, 3.0,
             2.2, 2.4, 3.5, 1.7, -1.3, -4.2, -6.0, -5.4, -1.5, 6.0, 12.6, 13.9, 12.3)


def _erb_centres(f_lo: float, f_hi: float):
    """ERB-spaced centre frequencies, 1 ERB apart, centred in [f_lo, f_hi]."""
    to_erb = lambda f: 9.2645 * np.sign(f) * np.log(1.0 + f * 0.00437)
    lo, hi = to_erb(np.float64(f_lo)), to_erb(np.float64(f_hi))
    span = hi - lo
    count = int(math.floor(span))
    pts = lo + np.arange(count + 1, dtype=np.float64) + (span - count) / 2.0
    centres = np.sign(pts) * (np.exp(np.abs(pts) / 9.2645) - 1.0) / 0.00437
    return centres, 24.7 + centres / 9.265


class MaskingModel:
    def __init__(self, block_size: int, sampling_rate: float, fullscale_db_spl: float = 94.0):
        nb, fs = int(block_size), float(sampling_rate)
        if nb % 2:
            raise RuntimeError("Block size is expected to be even")
        self.block_size, self.sampling_rate = nb, fs
        F = nb // 2 + 1
        f = np.arange(F, dtype=np.float64) * (fs / nb)
        p_full = 20e-6 * 10.0 ** (fullscale_db_spl / 20.0)
        thr_db = CubicSpline(np.array(ISO226_HZ, dtype=np.float64), np.array(ISO226_DB, dtype=np.float64),
                             bc_type="not-a-knot", extrapolate=True)(f)
        ear = p_full / (20e-6 * 10.0 ** (thr_db / 20.0))            # 1 / threshold in digital full scale
        cf, bw = _erb_centres(0.0, fs / 2.0)
        order = 4
        k = 2.0 ** (order - 1) * math.factorial(order - 1) / (math.pi * 15.0)   # (2*order-3)!! = 15
        bank = (1.0 + ((f[:, None] - cf[None, :]) / (k * bw[None, :])) ** 2) ** (-order / 2.0)   # (F, C)
        self.n_channels = bank.shape[1]
        self.G2 = np.ascontiguousarray(((ear[:, None] * bank) ** 2).T)          # (C, F)
        self.Leff = min(nb / fs / 0.3, 1.0)
        # calibration: detectability of a 52 dB probe in a 70 dB masker at the same bin equals 1
        k0 = max(nb // 48 - 1, 1)
        t = np.arange(nb, dtype=np.float64) / fs
        amp = lambda db: math.sqrt(2.0) * 20e-6 * 10.0 ** (db / 20.0) / p_full
        spec = lambda db: abs((math.sqrt(2.0) / nb * np.fft.fft(amp(db) * np.sin(2 * np.pi * f[k0] * t)))[k0])
        K = float(np.sum(bank[k0, :] ** 2) * self.Leff)
        g2 = self.G2[:, k0]
        k52, k70 = g2 * spec(52.0) ** 2, g2 * spec(70.0) ** 2
        fun = lambda x: self.Leff * float(np.sum(k52 / (k70 + x * K))) - 1.0 / x
        lo, hi = 0.1, 200.0
        if fun(hi) < 0:
            hi = 1000.0
        if np.sign(fun(lo)) == np.sign(fun(hi)):
            raise RuntimeError("Initialization of bisection method failed")
        mid = 0.5 * (lo + hi)
        for _ in range(1000):
            mid = 0.5 * (lo + hi)
            fm = fun(mid)
            done = fm == 0 or 0.5 * (hi - lo) < 1e-6
            if np.sign(fm) == np.sign(fun(lo)):
                lo = mid
            else:
                hi = mid
            if done:
                break
        self.Cs, self.Ca = float(mid), float(mid * K)

    def gain(self, time_block):
        """Host evaluation with the ``libdetectability`` call convention (``apvast.py:318``)."""
        x = np.asarray(time_block, dtype=np.float64).reshape(-1)
        p2 = (2.0 / self.block_size ** 2) * np.abs(np.fft.rfft(x)) ** 2
        pc = self.G2 @ p2
        return np.sqrt(self.Cs * self.Leff * np.sum(self.G2 / (pc[:, None] + self.Ca), axis=0))
