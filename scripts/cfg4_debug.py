import sys, os, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ap_vast_unofficial_b200 import _capi as capi, apvast
from ap_vast_unofficial_b200.workloads import make_workload
wl = make_workload("cfg3", n_blocks=8)
cfg = dict(wl["cfg"]); n = 4096; cfg["number_of_eigenvectors"] = n
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **cfg)
H = eng.hop_size
mus = np.logspace(-3, 1, 8)
d_w = torch.empty((8, 2, n, n), dtype=torch.float64, device="cuda")
lib = capi.lib()
for t in range(6):
    t0 = time.perf_counter()
    a = np.ascontiguousarray(wl["signal_A"][t*H:(t+1)*H]); b = np.ascontiguousarray(wl["signal_B"][t*H:(t+1)*H])
    da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
    capi.check(lib.apv_process_block_device(eng._h, C.c_void_p(da.data_ptr()), C.c_void_p(db.data_ptr())))
    capi.check(lib.apv_synchronize(eng._h)); t1 = time.perf_counter()
    met = eng.sweep_metrics(mus); torch.cuda.synchronize(); t1b = time.perf_counter()
    out = np.zeros(3)
    capi.check(lib.apv_sweep_device(eng._h, 8, capi.ptr(mus), C.c_void_p(d_w.data_ptr()), None)); t2 = time.perf_counter()
    lam = eng.lambda_A
    print("   metrics only %.1f ms, filters only %.1f ms; lambda min %.3e max %.3e nan %d; met nan %d" % (1e3*(t1b-t1), 1e3*(t2-t1b), lam.min(), lam.max(), int(np.isnan(lam).sum()), int(np.isnan(met).sum())))
    st = eng.stage_times()
    print("block %d: wall block %.1f ms, sweep %.1f ms; stages total %.1f S5 %.1f eig %.1f bt %.1f" % (t, 1e3*(t1-t0), 1e3*(t2-t1), st["total"], st["S5_jdiag"], st["S5_eig"], st["S5_backtransform"]), flush=True)
