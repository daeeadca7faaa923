"""TEST INFRASTRUCTURE ONLY -- golden vectors at BASELINE cfg-3 size (n = 4096) from the UNMODIFIED reference.

    python -m oracle.make_golden_cfg3 [nblocks]

Runs /root/reference/Python/apvast.py on the synthetic cfg-3 workload (ap_vast_unofficial_b200.workloads, which is
deterministic, so only the reference's results are stored): about one minute of CPU per block.  Stores, per block,
the leading eigenvalues, filters for a subset of ranks, r vectors, the diagonal and three rows of every R."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference  # noqa: E402
from ap_vast_unofficial_b200.workloads import make_workload  # noqa: E402


def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    ref = load_reference()
    wl = make_workload("cfg3", n_blocks=nblk)
    np.random.seed(0)
    ap = ref.apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
    H = ap.hop_size
    V = ap.number_of_eigenvectors
    ranks = [0, 1, 2, 3, 7, 15, 31, 47, 63]
    g = {"nblk": np.int64(nblk), "ranks": np.array(ranks), "seed": np.int64(0)}
    for t in range(nblk):
        t0 = time.time()
        outs = ap.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
        print("block", t, "%.1f s" % (time.time() - t0), flush=True)
        for z in ("A", "B"):
            g[f"w_{z}_{t}"] = getattr(ap, f"w_{z}")[ranks, :, 0].copy()
            g[f"lambda_{z}_{t}"] = getattr(ap, f"lambda_{z}")[:V + 1].copy()
            g[f"r_{z}_{t}"] = getattr(ap, f"r_{z}")[:, 0].copy()
        for nm in ("R_A_to_A", "R_A_to_B", "R_B_to_A", "R_B_to_B"):
            R = getattr(ap, nm)
            n = R.shape[0]
            g[f"{nm}_diag_{t}"] = np.diag(R).copy()
            g[f"{nm}_rows_{t}"] = R[[0, n // 2 - 1, n - 1], :].copy()
        g[f"out_A_{t}"] = np.stack([outs[0][v] for v in (0, V - 1)]).copy()
        g[f"out_B_{t}"] = np.stack([outs[1][v] for v in (0, V - 1)]).copy()
    path = os.path.join(ROOT, "tests", "golden", "cfg3_reference.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
