"""BASELINE cfg-5 shapes on one GPU: L = 32 loudspeakers, J = 256 taps (n = 8192), 4 zones of M microphones, one
clip; the 64 clips of the configuration are independent (8 per GPU on 8 GPUs, no communication).  Times the 4-zone
filter update (zones.apvast_zones: one two-zone engine per bright zone, dark zone = the other zones' microphones) and
checks the joint-diagonalisation identities of every zone.  Usage: python scripts/run_cfg5.py [n_blocks] [M] [L] [J]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200.zones import apvast_zones
from ap_vast_unofficial_b200.workloads import _rirs, _programme

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
J = int(sys.argv[4]) if len(sys.argv) > 4 else 256
Z, K, Nb, N, V = 4, 1024, 2048, 2048, 64
H = Nb // 2
rirs = [_rirs(20 + z, K, L, M) for z in range(Z)]
sigs = [_programme(40 + z, nblk * H) for z in range(Z)]
np.random.seed(0)
t0 = time.time()
eng = apvast_zones(Nb, rirs, J, 32, [0, 1, 2, 3], V, 1.0, N)
print("4 zones, L=%d J=%d n=%d, M=%d per zone: engines created in %.1f s" % (L, J, L * J, M, time.time() - t0), flush=True)
for t in range(nblk):
    t0 = time.time()
    outs = eng.process_input_buffers([s[t * H:(t + 1) * H] for s in sigs])
    dt = time.time() - t0
    st = eng.stage_times()
    print("block %d: %.2f s wall for the 4 zones | per zone S4 %.0f ms, S5 %.0f ms (tridiag %.0f), total %.0f ms" %
          (t, dt, np.mean([s["S4_stats"] for s in st]), np.mean([s["S5_jdiag"] for s in st]),
           np.mean([s["S5_tridiag"] for s in st]), np.mean([s["total"] for s in st])), flush=True)
n = L * J
for z in range(Z):
    RB, RD, rB = eng.statistics(z)
    U = np.array(eng.engines[z].U_A); lam = np.array(eng.eigenvalues[z])
    i1 = np.max(np.abs(U.T @ (RD + 1e-7 * np.eye(n)) @ U - np.eye(V)))
    i2 = np.max(np.abs(U.T @ RB @ U - np.diag(lam))) / lam[0]
    print("zone %d: |U'(R_D + reg I)U - I| = %.1e, |U'R_B U - Lambda| / lambda_1 = %.1e" % (z, i1, i2), flush=True)
print("updates/s for 4-zone blocks on this GPU: %.3f  (8 GPUs, clips independent: x8)" % (1.0 / dt))
