import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload
wl = make_workload("cfg2", n_blocks=5)
res = {}
for mode in (1, 0):
    np.random.seed(0)
    e = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, run_B=False, eig_mode=mode, **wl["cfg"])
    H = e.hop_size
    for t in range(5):
        o = e.process_input_buffers(wl["signal_A"][t*H:(t+1)*H], wl["signal_B"][t*H:(t+1)*H])
    res[mode] = (np.array(e.w_A), np.array(o[0]))
    assert o[1] is None and e.w_B is None
    print("mode", mode, {k: round(v, 2) for k, v in e.stage_times().items() if k.startswith("S5") or k == "total"})
    e.close()
V = res[0][0].shape[0]
d = np.linalg.norm((res[0][0] - res[1][0]).reshape(V, -1), axis=1) / np.linalg.norm(res[1][0].reshape(V, -1), axis=1)
print("one zone (run_B=False), n=1024: auto (two-stage) vs one-stage filters %.2e, outputs %.2e" % (d.max(), np.linalg.norm(res[0][1] - res[1][1]) / np.linalg.norm(res[1][1])))
