"""Soak: N blocks of a workload through the pipelined multi-hop call against N per-hop calls, compared BITWISE (outputs
and filters of every block), twice.  A race in one of the hand-synchronised kernels (bulge chasing, Q2 wavefront,
cluster QR) or a missing dependency between the two streams shows up as a bit difference.
    python scripts/soak_pipeline.py [workload] [nblocks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.workloads import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = make_workload(name, n_blocks=nb)
H = wl["hop"]
np.random.seed(0); e1 = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
ref_o, ref_w = [], []
t0 = time.time()
for t in range(nb):
    oA, oB, _, _ = e1.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
    ref_o.append((np.stack(oA), np.stack(oB))); ref_w.append((e1.w_A[:, :, 0].copy(), e1.w_B[:, :, 0].copy()))
t_seq = time.time() - t0
bad = 0
for rep in range(2):
    np.random.seed(0); e2 = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
    t0 = time.time()
    oA, oB, _, _, w = e2.process_blocks(wl["signal_A"], wl["signal_B"], want_filters=True)
    t_pipe = time.time() - t0
    for t in range(nb):
        ok = (np.array_equal(oA[t], ref_o[t][0]) and np.array_equal(oB[t], ref_o[t][1]) and
              np.array_equal(w[t, 0], ref_w[t][0]) and np.array_equal(w[t, 1], ref_w[t][1]))
        bad += 0 if ok else 1
    print(f"repeat {rep}: {nb} blocks of {name}: per-hop loop {t_seq:.2f} s, process_blocks {t_pipe:.2f} s, "
          f"blocks with any bit difference so far: {bad}", flush=True)
    e2.close()
assert bad == 0
print("OK: bitwise identical")
