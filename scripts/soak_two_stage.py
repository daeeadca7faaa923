"""Soak test of the hand-synchronised kernels of the two-stage route (cluster QR, chase wavefront, chained
back-transformations): many blocks, filters of the automatic route against the one-stage route on every block."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

for name, nblk in ((sys.argv[1], int(sys.argv[2])),) if len(sys.argv) > 2 else (("cfg2", 150), ("cfg3", 40)):
    wl = make_workload(name, n_blocks=nblk)
    np.random.seed(0); e3 = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, eig_mode=0, **wl["cfg"])
    np.random.seed(0); e1 = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, eig_mode=1, **wl["cfg"])
    H, V = e3.hop_size, e3.number_of_eigenvectors
    worst = 0.0; t0 = time.time()
    for t in range(nblk):
        a, b = wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H]
        o3 = e3.process_input_buffers(a, b); o1 = e1.process_input_buffers(a, b)
        for zn in ("A", "B"):
            w3, w1 = np.array(getattr(e3, "w_" + zn))[:, :, 0], np.array(getattr(e1, "w_" + zn))[:, :, 0]
            lam = np.array(getattr(e1, "lambda_" + zn))
            gap = np.abs(np.diff(lam)) / lam[0]
            for v in range(V - 1):
                if gap[v] > 1e-9:
                    worst = max(worst, float(np.linalg.norm(w3[v] - w1[v]) / np.linalg.norm(w1[v])))
        eo = float(np.linalg.norm(np.array(o3[0]) - np.array(o1[0])) / np.linalg.norm(np.array(o1[0])))
        worst = max(worst, eo)
        assert worst < 1e-8, (name, t, worst)
    print("%s: %d blocks, worst relative difference (filters of resolved ranks, rendered outputs) two-stage vs one-stage: %.2e  (%.1f s)"
          % (name, nblk, worst, time.time() - t0), flush=True)
    e3.close(); e1.close()
