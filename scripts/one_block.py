"""Run a few blocks of a named workload (for ncu captures): python scripts/one_block.py cfg3 [nblocks]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = make_workload(name, n_blocks=nb)
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
H = eng.hop_size
for t in range(nb):
    eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
print(eng.stage_times())
