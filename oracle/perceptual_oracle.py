"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's in-repo perceptual model.

PARITY UNPINNED.  The Python reference obtains its perceptual gain from the third-party package
``libdetectability`` (unpinned version; imported at ``Python/apvast.py:4``, constructed at
``:77-83``, called at ``:318-319``).  That package is not under ``/root/reference`` and is not
installed, and the reference holds no golden vector at that boundary.  What *is* in the reference is
the MATLAB twin of the same van-de-Par (2005) spectral-integration model, restated here:

* ``Matlab/ControlMethods/perceptualModel.m:30-116``   tables + calibration (bisection)
* ``Matlab/ControlMethods/perceptualModel.m:118-139``  squared weighting curve
* ``Matlab/ControlMethods/gammatoneFilterResponse.m:11-19,32-51``  gammatone magnitude bank
* ``Matlab/ControlMethods/interpolatedThresholdOfHearing.m:19-20,29-30``  ISO 226:2003 threshold,
  cubic spline *with extrapolation* (MATLAB ``interp1(...,'spline')`` = not-a-knot spline).

The object exposes the Python call convention of ``Python/apvast.py:318``:
``gain(time_block) -> (Nb/2+1,)`` real array.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.interpolate import CubicSpline

# interpolatedThresholdOfHearing.m:29-30
_ISO_F = np.array([20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800,
                   1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500], dtype=np.float64)
_ISO_SPL = np.array([78.5, 68.7, 59.5, 51.1, 44.0, 37.5, 31.5, 26.5, 22.1, 17.9, 14.4, 11.4, 8.6, 6.2,
                     4.4, 3.0, 2.2, 2.4, 3.5, 1.7, -1.3, -4.2, -6.0, -5.4, -1.5, 6.0, 12.6, 13.9, 12.3],
                    dtype=np.float64)


def threshold_of_hearing_db(frequency):
    """interpolatedThresholdOfHearing.m:19-20 (default method iso226_2003)."""
    cs = CubicSpline(_ISO_F, _ISO_SPL, bc_type="not-a-knot", extrapolate=True)
    return cs(np.asarray(frequency, dtype=np.float64))


def gammatone_response(flow, fhigh, frequency):
    """gammatoneFilterResponse.m:8-19 with helper :32-51.  Returns (F, C)."""
    order = 4
    lim = np.array([flow, fhigh], dtype=np.float64)
    erb_lim = 9.2645 * np.sign(lim) * np.log(1.0 + lim * 0.00437)
    erb_range = erb_lim[1] - erb_lim[0]
    n = int(math.floor(erb_range / 1.0))
    rem = erb_range - n * 1.0
    erb_pts = erb_lim[0] + np.arange(n + 1, dtype=np.float64) * 1.0 + rem / 2.0
    cf = (1.0 / 0.00437) * np.sign(erb_pts) * (np.exp(np.abs(erb_pts) / 9.2645) - 1.0)
    bw = 24.7 + cf / 9.265
    # k = 2^(o-1) (o-1)! / (pi (2o-3)!!)  ; (2*4-3)!! = 5!! = 15
    dfact = 1.0
    for t in range(1, 2 * order - 3 + 1, 2):
        dfact *= t
    k = 2.0 ** (order - 1) * math.factorial(order - 1) / (math.pi * dfact)
    f = np.asarray(frequency, dtype=np.float64).reshape(-1, 1)
    return (1.0 + ((f - cf.reshape(1, -1)) / (k * bw.reshape(1, -1))) ** 2) ** (-order / 2.0)


class PerceptualModelOracle:
    """perceptualModel.m restated; ``fullscale_db`` = MATLAB ``fullscalePressureInDbSpl``."""

    def __init__(self, block_size: int, sampling_rate: float, fullscale_db: float = 94.0):
        if block_size % 2 != 0:
            raise RuntimeError("Block size is expected to be even")
        nb = int(block_size)
        fs = float(sampling_rate)
        self.block_size = nb
        self.sampling_rate = fs
        fullscale_pa = 10.0 ** (fullscale_db / 20.0) * 20e-6
        nf = nb // 2 + 1
        freq = np.arange(nf, dtype=np.float64) * (fs / nb)
        thr_db = threshold_of_hearing_db(freq)
        thr_pa = 10.0 ** (thr_db / 20.0) * 20e-6
        thr = thr_pa / fullscale_pa
        self.ome = 1.0 / thr                                            # :46
        self.fb = gammatone_response(0.0, fs / 2.0, freq)               # :49
        self.n_channels = self.fb.shape[1]
        self.G = self.ome.reshape(-1, 1) * self.fb                      # :51-53
        self.Leff = min(nb / fs / 0.3, 1.0)                             # :56
        a52 = math.sqrt(2.0) * 10.0 ** (52 / 20.0) * 20e-6 / fullscale_pa
        a70 = math.sqrt(2.0) * 10.0 ** (70 / 20.0) * 20e-6 / fullscale_pa
        f_idx = nb // 48                    # MATLAB 1-based index -> 0-based bin f_idx-1 (:66-67,75-76)
        k0 = max(f_idx - 1, 1)              # guard for tiny test sizes (MATLAB would index 0)
        cal_f = freq[k0]
        t = np.arange(nb, dtype=np.float64) / fs
        s52 = math.sqrt(2.0) / nb * np.fft.fft(a52 * np.sin(2 * np.pi * cal_f * t))
        s70 = math.sqrt(2.0) / nb * np.fft.fft(a70 * np.sin(2 * np.pi * cal_f * t))
        S52 = abs(s52[k0])
        S70 = abs(s70[k0])
        K = float(np.sum(self.fb[k0, :] ** 2) * self.Leff)              # :78
        k52 = self.G[k0, :] ** 2 * S52 ** 2
        k70 = self.G[k0, :] ** 2 * S70 ** 2

        def fun(x):
            return self.Leff * np.sum(k52 / (k70 + x * K)) - 1.0 / x

        x_neg, x_pos = 1e-1, 200.0
        if fun(x_pos) < 0:
            x_pos = 1000.0
        if np.sign(fun(x_neg)) == np.sign(fun(x_pos)):
            raise RuntimeError("Initialization of bisection method failed")
        itr, found, x_mid = 1, False, None
        while itr < 1000 and not found:                                 # :94-106
            x_mid = (x_pos + x_neg) / 2.0
            f_mid = fun(x_mid)
            if f_mid == 0 or (x_pos - x_neg) / 2.0 < 1e-6:
                found = True
            itr += 1
            if np.sign(f_mid) == np.sign(fun(x_neg)):
                x_neg = x_mid
            else:
                x_pos = x_mid
        self.Cs = float(x_mid)
        self.Ca = float(x_mid * K)

    def squared_weighting_curve(self, time_block):
        """perceptualModel.m:118-139 for a real time block."""
        nb = self.block_size
        x = np.asarray(time_block, dtype=np.float64).reshape(-1)
        if x.size != nb:
            raise RuntimeError("The size of the inputBlock does not match the expected blockSize")
        spec = math.sqrt(2.0) / nb * np.fft.rfft(x)
        mag = np.abs(spec).reshape(-1, 1)
        masker = np.sum((self.G * mag) ** 2, axis=0)                    # (C,)
        return self.Cs * self.Leff * np.sum(self.G ** 2 / (masker.reshape(1, -1) + self.Ca), axis=1)

    def gain(self, time_block):
        """Python call convention (``apvast.py:318``): F real gains = sqrt(weighting curve^2)."""
        return np.sqrt(self.squared_weighting_curve(time_block))
