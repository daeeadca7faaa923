"""TEST INFRASTRUCTURE ONLY -- rendered loudspeaker feeds of the UNMODIFIED reference at cfg-1 (Python/rirs.mat, the
make_python_test.m:6-15 parameters) over a run of hops, and the acoustic contrast / normalised signal distortion of
those feeds at the control microphones of rirs.mat (definitions: Matlab/main.m:120-130, predictPressure.m:12-17).

    python -m oracle.make_golden_cfg1_feeds        ->  tests/golden/cfg1_feeds.npz

The north-star criterion "acoustic contrast and normalised signal distortion within 0.01 dB" is checked against these
numbers by tests/test_gpu_metrics_sharded.py::test_cfg1_contrast_and_distortion_against_reference_feeds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference, reference_rirs  # noqa: E402
from ap_vast_unofficial_b200.metrics import evaluate_zone  # noqa: E402


def main():
    nblk, skip = 16, 6           # metrics over the hops after the start transient (random start buffers flushed)
    ref = load_reference()
    rA, rB = reference_rirs()
    cfg = dict(block_size=1600, filter_length=100, modeling_delay=20, reference_index_A=6, reference_index_B=6,
               number_of_eigenvectors=50, mu=1.0, statistics_buffer_length=1000)
    np.random.seed(0)
    ap = ref.apvast(rir_A=rA, rir_B=rB, perceptual=False, **cfg)
    H = ap.hop_size
    rng = np.random.default_rng(1)
    sA, sB = rng.standard_normal(nblk * H), rng.standard_normal(nblk * H)
    ranks = [0, 24, 49]
    fa, fb = [], []
    for t in range(nblk):
        oA, oB, _, _ = ap.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        fa.append(np.stack([oA[v] for v in ranks]).copy())
        fb.append(np.stack([oB[v] for v in ranks]).copy())
        print("hop", t, flush=True)
    fa, fb = np.concatenate(fa, axis=1), np.concatenate(fb, axis=1)          # (ranks, T, L)
    met = np.zeros((len(ranks), 2, 2))
    for i in range(len(ranks)):
        for z, (f, rb, rd, sig, refidx) in enumerate(((fa[i], rA, rB, sA, 6), (fb[i], rB, rA, sB, 6))):
            met[i, z] = evaluate_zone(f[skip * H:], rb, rd, sig[skip * H:], refidx, 20)
    g = dict(input_A=sA, input_B=sB, ranks=np.array(ranks), nblk=np.int64(nblk), skip=np.int64(skip), seed=np.int64(0),
             feeds_A=fa, feeds_B=fb, metrics=met)
    for k, v in cfg.items():
        g["cfg_" + k] = np.array(v)
    path = os.path.join(ROOT, "tests", "golden", "cfg1_feeds.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB; AC/NSD [rank][zone]:\n", met)


if __name__ == "__main__":
    main()
