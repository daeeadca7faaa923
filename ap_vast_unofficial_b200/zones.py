"""More than two sound zones (BASELINE.json configs[4]) on top of the two-zone block
engine.  The reference (``Python/apvast.py``) has exactly two zones; the generalisation used here is the one SURVEY
8(d) states: for bright zone z the dark-zone statistics are the sum over the other zones,

    R_B(z) = R_{z -> z},      R_D(z) = sum_{z' != z} R_{z -> z'},      r_B(z) = r_{z -> z},

followed by the same joint diagonalisation and rank-V filter sum.  Because R_{z -> z'} is a sum over the microphones of
zone z' (``apvast.py:332-364`` loops over m), R_D(z) is exactly the cross-zone matrix of a TWO-zone problem whose
second zone holds the microphones of all the other zones.  One two-zone engine per bright zone therefore does the
whole job on the GPU, unchanged: zone A = zone z (padded with silent microphones to the common count, which the
statistics skip: ``active_mics_A``), zone B = the union of the other zones, ``run_B=False``.  The loudspeaker signal of the array is the sum of the zones' feeds.

Per-zone perceptual weighting (``perceptual=True``): every microphone is weighted with the masking curve derived from
the target signal of ITS OWN zone -- the generalisation of ``apvast.py:259-262,318-319``, where zone-A microphones use
W_A (from target A->A) and zone-B microphones W_B (from target B->B).  Engine z computes the curves of its bright
microphones in S2; the dark microphones of engine z that belong to zone q take engine q's curves.  The engines
therefore run the block in two halves (``perceptual`` mode 3 of the C-ABI): S1 + S2 on all engines, a device-to-device
exchange of the weighting curves between the engines (``apv_copy_weights``), then S3..S7.
"""
from __future__ import annotations

import numpy as np

from .apvast import apvast


class apvast_zones:
    """Z-zone AP-VAST: ``rirs`` is a list of Z arrays (K, L, M); ``reference_indices`` one loudspeaker per zone.

    ``process_input_buffers(inputs)`` takes Z hops and returns a list of Z lists of V arrays (H, L) (the feeds that
    render programme z), like ``output_buffer_A`` of the two-zone call.  ``w[z]`` are the filters (V, n, 1)."""

    def __init__(self, block_size, rirs, filter_length, modeling_delay, reference_indices, number_of_eigenvectors, mu,
                 statistics_buffer_length, hop_size=None, sampling_rate=48000, **engine_kwargs):
        rirs = [np.ascontiguousarray(r, dtype=np.float64) for r in rirs]
        if len(rirs) < 2:
            raise RuntimeError("at least two zones")
        if any(r.shape != rirs[0].shape for r in rirs):
            raise RuntimeError("rirs of unequal size")
        self.perceptual = bool(engine_kwargs.pop("perceptual", False))
        concurrent = bool(engine_kwargs.pop("concurrent", True))
        K, L, M = rirs[0].shape
        Z = len(rirs)
        self.n_zones, self.number_of_eigenvectors = Z, int(number_of_eigenvectors)
        self.engines = []
        for z in range(Z):
            bright = rirs[z] if Z == 2 else np.concatenate([rirs[z], np.zeros((K, L, M * (Z - 2)))], axis=2)
            dark = np.concatenate([rirs[q] for q in range(Z) if q != z], axis=2)
            self.engines.append(apvast(block_size, bright, dark, filter_length, modeling_delay, reference_indices[z], 0,
                                       number_of_eigenvectors, mu, statistics_buffer_length, hop_size, sampling_rate,
                                       run_A=True, run_B=False, perceptual=self.perceptual, split_weights=self.perceptual,
                                       active_mics_A=M, **engine_kwargs))
        self.hop_size = self.engines[0].hop_size
        self._silence = np.zeros(self.hop_size)
        self._M = M
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=Z if concurrent else 1)

    def _exchange_weights(self):
        from . import _capi as capi
        lib = capi.lib()
        Z, M = self.n_zones, self._M
        for z in range(Z):
            others = [q for q in range(Z) if q != z]
            for i, q in enumerate(others):      # dark microphones [i M, (i+1) M) of engine z are zone q's microphones
                capi.check(lib.apv_copy_weights(self.engines[z]._h, 1, i * M, self.engines[q]._h, 0, 0, M))

    def process_input_buffers(self, inputs):
        if len(inputs) != self.n_zones:
            raise RuntimeError("invalid input size")
        # The Z engines are independent and each owns its CUDA streams: one host thread per engine (ctypes releases the
        # GIL) lets the latency-bound phases of one zone's joint diagonalisation (bulge chasing, eigenvectors) run
        # beside the tensor-core phases of the others instead of one zone after the other.
        if not self.perceptual:
            return list(self._pool.map(lambda ex: ex[0].process_input_buffers(ex[1], self._silence)[0],
                                       zip(self.engines, inputs)))
        list(self._pool.map(lambda ex: ex[0]._begin(ex[1], self._silence), zip(self.engines, inputs)))
        self._exchange_weights()
        return list(self._pool.map(lambda e: e._finish()[0], self.engines))

    @property
    def w(self):
        return [eng.w_A for eng in self.engines]

    @property
    def eigenvalues(self):
        return [eng.lambda_A for eng in self.engines]

    def statistics(self, z):
        """(R_B, R_D, r_B) of bright zone z as the engine holds them."""
        e = self.engines[z]
        return e.R_A_to_A, e.R_A_to_B, e.r_A

    def stage_times(self):
        return [eng.stage_times() for eng in self.engines]

    def close(self):
        self._pool.shutdown(wait=True)
        for e in self.engines:
            e.close()
