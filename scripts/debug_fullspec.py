"""Diagnostic: B-orthonormality of the full set of joint eigenvectors at cfg-2 size (V = n = 1024), by index range."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload
wl = make_workload("cfg2", n_blocks=5)
cfg = dict(wl["cfg"]); n = 1024
cfg["number_of_eigenvectors"] = n
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **cfg)
H = eng.hop_size
for t in range(5):
    eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
if os.environ.get("DBG_SWEEP"):
    mus = np.array([0.01, 1.0, 10.0])
    wA, wB = eng.sweep(mus)
    met = eng.sweep_metrics(mus)
RB, RD, U, lam = (eng.R_B_to_B, eng.R_B_to_A, eng.U_B, eng.lambda_B) if os.environ.get('DBG_ZONE_B') else (eng.R_A_to_A, eng.R_A_to_B, eng.U_A, eng.lambda_A)
Bm = RD + 1e-7 * np.eye(n)
E = np.abs(U.T @ Bm @ U - np.eye(n))
print("env", {k: v for k, v in os.environ.items() if k.startswith("APV_")})
print("lambda head/tail", lam[:3], lam[-5:], "min rel gap", np.min(np.abs(np.diff(lam))) / lam[0])
for a in range(0, n, 128):
    print("idx %4d..%4d  max |U'BU - I| diag-block %.2e  vs-all %.2e" % (a, a + 127, E[a:a + 128, a:a + 128].max(), E[a:a + 128, :].max()))
A = U.T @ RB @ U
print("max |U'AU - L| / l0", np.max(np.abs(A - np.diag(lam))) / lam[0])
import scipy.linalg as sla
lref = sla.eigh(RB, Bm, eigvals_only=True)[::-1]
print("eigenvalue error vs scipy eigh", np.max(np.abs(lam - lref)) / lref[0])
print(eng.stage_times())
