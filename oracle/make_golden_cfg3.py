"""TEST INFRASTRUCTURE ONLY -- golden vectors at BASELINE cfg-2 / cfg-3 size from the UNMODIFIED reference.

    python -m oracle.make_golden_cfg3 [nblocks]                       # tests/golden/cfg3_reference.npz (round 1)
    python -m oracle.make_golden_cfg3 --workload cfg3 --variant 1 --nblocks 8 --all-ranks-from 4
                                                                     # tests/golden/cfg3_reference_v1.npz
    python -m oracle.make_golden_cfg3 --workload cfg2 --nblocks 10 --all-ranks-from 0
                                                                     # tests/golden/cfg2_reference.npz

Runs /root/reference/Python/apvast.py on the synthetic workload (ap_vast_unofficial_b200.workloads, which is
deterministic, so only the reference's results are stored): about one minute of CPU per block at cfg-3, 4 s at
cfg-2.  Stores, per block, the leading eigenvalues, the filters (a subset of ranks, or ALL ranks from block
`--all-ranks-from` on -- the blocks past the warm-up of the statistics buffers), the r vectors, the diagonal and
three rows of every R, and the rendered outputs of the first and the last rank."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference  # noqa: E402
from ap_vast_unofficial_b200.workloads import make_workload  # noqa: E402


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("nblocks_pos", nargs="?", type=int, default=None)
    ap_.add_argument("--workload", default="cfg3")
    ap_.add_argument("--variant", type=int, default=0)
    ap_.add_argument("--nblocks", type=int, default=5)
    ap_.add_argument("--all-ranks-from", type=int, default=-1, help="store all V ranks from this block on (-1: never)")
    ap_.add_argument("--seed", type=int, default=0)
    a = ap_.parse_args()
    nblk = a.nblocks_pos if a.nblocks_pos is not None else a.nblocks
    ref = load_reference()
    wl = make_workload(a.workload, n_blocks=nblk, variant=a.variant)
    np.random.seed(a.seed)
    ap = ref.apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
    H = ap.hop_size
    V = ap.number_of_eigenvectors
    sub = [r for r in (0, 1, 2, 3, 7, 15, 31, 47, 63) if r < V]
    g = {"nblk": np.int64(nblk), "ranks": np.array(sub), "seed": np.int64(a.seed), "variant": np.int64(a.variant),
         "all_ranks_from": np.int64(a.all_ranks_from)}
    for t in range(nblk):
        t0 = time.time()
        outs = ap.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
        print("block", t, "%.1f s" % (time.time() - t0), flush=True)
        ranks = list(range(V)) if 0 <= a.all_ranks_from <= t else sub
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
    name = f"{a.workload}_reference" + (f"_v{a.variant}" if a.variant else "") + ".npz"
    path = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
