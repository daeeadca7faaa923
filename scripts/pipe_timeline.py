"""Timeline of the pipelined multi-block path at cfg-3 (diagnostic): when do the front halves (S1-S4, low-priority
stream) run relative to the back halves (S5-S7)?   python scripts/pipe_timeline.py [nblocks] [workload]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ap_vast_unofficial_b200 import _capi as capi, apvast  # noqa: E402
from ap_vast_unofficial_b200.sharded import RangeRunner  # noqa: E402
from ap_vast_unofficial_b200.workloads import make_workload  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
wl = make_workload(sys.argv[2] if len(sys.argv) > 2 else "cfg3", n_blocks=nb)
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
lib = capi.lib()
rr = RangeRunner(eng, 0, 1, None, max_owned=nb)
H = eng.hop_size
rr.run(wl["signal_A"], wl["signal_B"], 0, nb); rr.gather([nb], None, None)       # warm
capi.check(lib.apv_debug_timeline(eng._h, nb, None))
rr.run(wl["signal_A"], wl["signal_B"], 0, nb); rr.gather([nb], None, None)
ms = (C.c_float * (4 * nb))()
capi.check(lib.apv_debug_timeline(eng._h, nb, ms))
t = np.array(ms[:]).reshape(nb, 4)
print("block  front_start  front_end  back_start  back_end   (ms)")
for b in range(nb):
    print("%5d  %10.2f  %9.2f  %10.2f  %8.2f" % (b, *t[b]))
print("stage times of the last block:", eng.stage_times())
