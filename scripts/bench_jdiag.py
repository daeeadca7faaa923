"""Per-stage device time of one block by eigen-solver mode (1 = one-stage tridiagonalisation, 3 = two-stage) and
the filter difference between the modes.  Usage: python scripts/bench_jdiag.py [cfg2|cfg3|cfg5_2zone ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

for name in (sys.argv[1:] or ["cfg3"]):
    wl = make_workload(name, n_blocks=6)
    ws = {}
    for mode in (1, 3):
        np.random.seed(0)
        eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, eig_mode=mode, **wl["cfg"])
        H = eng.hop_size
        for t in range(5):
            eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
        st = eng.stage_times()
        print(name, "eig_mode", mode, {k: round(v, 2) for k, v in st.items()}, flush=True)
        ws[mode] = (np.array(eng.w_A), np.array(eng.lambda_A))
        eng.close()
    V = ws[1][0].shape[0]
    d = np.linalg.norm((ws[1][0] - ws[3][0]).reshape(V, -1), axis=1) / np.linalg.norm(ws[1][0].reshape(V, -1), axis=1)
    print(name, "filters mode 3 vs 1: max rel L2 over ranks %.3e; eigenvalues %.3e" %
          (d.max(), np.max(np.abs(ws[1][1] - ws[3][1])) / ws[1][1][0]), flush=True)
