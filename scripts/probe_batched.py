"""Diagnostic: wall time of apvast.process_blocks (K hops per call) after a warm call.  python scripts/probe_batched.py [K ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

Ks = [int(a) for a in sys.argv[1:]] or [8, 16]
wl = make_workload("cfg3", n_blocks=max(Ks) + 2)
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
H = eng.hop_size
a, b = wl["signal_A"], wl["signal_B"]
eng.process_blocks(a[:2 * H], b[:2 * H])
for K in Ks:
    for rep in range(2):
        t0 = time.perf_counter()
        eng.process_blocks(a[:K * H], b[:K * H], want_filters=True)
        dt = time.perf_counter() - t0
        print("process_blocks K=%d: %.1f ms/step (%.0f ms)" % (K, 1e3 * dt / K, 1e3 * dt), flush=True)
