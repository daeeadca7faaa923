"""Run a few blocks of a named workload (for ncu captures): python scripts/one_block.py cfg3 [nblocks].
Environment: APV_OB_EIG_MODE (eig_mode), APV_OB_V (number_of_eigenvectors, 0 = full rank), APV_OB_STATS (stats_mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = make_workload(name, n_blocks=nb)
np.random.seed(0)
cfg = dict(wl["cfg"])
if "APV_OB_V" in os.environ:
    v = int(os.environ["APV_OB_V"])
    cfg["number_of_eigenvectors"] = v if v > 0 else wl["shapes"]["n"]
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, eig_mode=int(os.environ.get("APV_OB_EIG_MODE", "0")),
             stats_mode=int(os.environ.get("APV_OB_STATS", "0")), **cfg)
H = eng.hop_size
for t in range(nb):
    eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
print(eng.stage_times())
