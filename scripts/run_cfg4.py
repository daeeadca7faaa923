"""BASELINE cfg-4: mu x V trade-off sweep (8 mu values x the full rank range) at L=16, J=256 (n=4096), one joint
diagonalisation per zone and block.  Prints device times and checks the full-rank filters against the closed form
w = (R_B + mu (R_D + reg I))^-1 r_B (apVast.m:115-118).  Usage: python scripts/run_cfg4.py [V]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = make_workload("cfg3", n_blocks=4)
cfg = dict(wl["cfg"], number_of_eigenvectors=V)
np.random.seed(0)
t0 = time.time()
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, render=False, **cfg) if "render" in apvast.__init__.__code__.co_varnames else \
    apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **cfg)
print("create %.1f s" % (time.time() - t0), flush=True)
H = eng.hop_size
for t in range(3):
    t0 = time.time()
    eng.advance_state(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H]) if hasattr(eng, "advance_state") else \
        eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
    print("warm-up block %d: %.2f s wall" % (t, time.time() - t0), flush=True)
t0 = time.time()
eng.process_input_buffers(wl["signal_A"][3 * H:4 * H], wl["signal_B"][3 * H:4 * H])
print("full block (V = %d): %.2f s wall" % (V, time.time() - t0), {k: round(v, 1) for k, v in eng.stage_times().items()}, flush=True)
mus = np.logspace(-3, 1, 8)
t0 = time.time()
wA, wB = eng.sweep(mus)
print("sweep of 8 mu x %d ranks: %.2f s wall; output %s" % (V, time.time() - t0, wA.shape), flush=True)
n = wA.shape[-1]
RA, RD, r = np.array(eng.R_A_to_A), np.array(eng.R_A_to_B), np.array(eng.r_A)[:, 0]
for k in (0, 4, 7):
    closed = np.linalg.solve(RA + mus[k] * (RD + 1e-7 * np.eye(n)), r)
    err = np.linalg.norm(wA[k, -1] - closed) / np.linalg.norm(closed)
    print("mu = %.3g: full-rank filter vs closed form: rel L2 %.2e" % (mus[k], err), flush=True)
