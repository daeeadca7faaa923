"""Diagnostic: per-call wall time of the drop-in call after a device-resident warm-up (the sequence bench.py uses),
with and without the nvidia-smi clock sampler of bench.py running."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ap_vast_unofficial_b200 import apvast, _capi as capi
from ap_vast_unofficial_b200.workloads import make_workload
import bench

wl = make_workload("cfg3", n_blocks=40)
np.random.seed(0)
eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
H = eng.hop_size
lib = capi.lib()
d_sig = torch.from_numpy(np.stack([wl["signal_A"], wl["signal_B"]])).cuda()
def dev_ptr(sig, blk): return C.c_void_p(d_sig.data_ptr() + (sig * d_sig.shape[1] + blk * H) * 8)
def blk(t): return wl["signal_A"][t*H:(t+1)*H], wl["signal_B"][t*H:(t+1)*H]
t = 0
for _ in range(4):
    capi.check(lib.apv_process_block_device(eng._h, dev_ptr(0, t), dev_ptr(1, t))); t += 1
capi.check(lib.apv_synchronize(eng._h))
for label, sampler in (("no sampler", None), ("with nvidia-smi sampler", bench.ClockSampler(0))):
    if sampler: sampler.start(); time.sleep(0.5)
    capi.check(lib.apv_timer_start(eng._h))
    for _ in range(6):
        capi.check(lib.apv_process_block_device(eng._h, dev_ptr(0, t), dev_ptr(1, t))); t += 1
    ms = C.c_float(0); capi.check(lib.apv_timer_stop(eng._h, C.byref(ms)))
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); eng.process_input_buffers(*blk(t)); ts.append(1e3 * (time.perf_counter() - t0)); t += 1
    print("%-26s device-resident %.1f ms/step | drop-in calls ms: %s" % (label, ms.value / 6, " ".join("%.1f" % x for x in ts)), flush=True)
    if sampler: print("   clocks", sampler.stop())
