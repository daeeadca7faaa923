"""CPU: the I/O helpers around the engine -- .mat RIR files, WAV feeders, hop streaming."""
import os

import numpy as np
import pytest

from ap_vast_unofficial_b200 import io as apio


def test_load_rirs_mat_roundtrip(tmp_path):
    from scipy.io import savemat
    rng = np.random.default_rng(0)
    rA, rB = rng.standard_normal((20, 3, 4)), rng.standard_normal((20, 3, 4))
    p = str(tmp_path / "rirs.mat")
    savemat(p, {"rirA": rA, "rirB": rB})
    a, b = apio.load_rirs_mat(p)
    assert a.flags.c_contiguous and a.dtype == np.float64
    assert np.array_equal(a, rA) and np.array_equal(b, rB)
    savemat(p, {"rirA": rA, "rirB": rB[:10]})
    with pytest.raises(RuntimeError, match="rirs of unequal size"):
        apio.load_rirs_mat(p)
    savemat(p, {"x": rA})
    with pytest.raises(KeyError):
        apio.load_rirs_mat(p)


@pytest.mark.skipif(not os.path.isfile("/root/reference/Python/rirs.mat"), reason="reference checkout not present")
def test_reference_rirs_mat_matches_golden_fixture():
    """The fixture of the reference (Python/rirs.mat) loads to the arrays stored with the cfg-1 golden vectors."""
    from tests._golden import load_case
    g, cfg, ctor = load_case("cfg1")
    a, b = apio.load_rirs_mat("/root/reference/Python/rirs.mat")
    assert np.array_equal(a, g["rir_A"]) and np.array_equal(b, g["rir_B"])


def test_wav_roundtrip_and_hops(tmp_path):
    from scipy.io import wavfile
    fs = 8000
    x = np.sin(2 * np.pi * 440 * np.arange(1000) / fs)
    p16, pf = str(tmp_path / "a16.wav"), str(tmp_path / "af.wav")
    wavfile.write(p16, fs, (x * 32767).astype(np.int16))
    fs2, y = apio.read_wav_mono(p16)
    assert fs2 == fs and np.max(np.abs(y - x)) < 1e-4
    apio.write_wav(pf, fs, np.stack([x, -x], axis=1))
    fs3, z = apio.read_wav_mono(pf)                # two channels average to zero
    assert fs3 == fs and np.max(np.abs(z)) < 1e-7
    blocks = list(apio.hop_blocks(x, x[:300], 256))
    assert len(blocks) == 4 and all(a.size == 256 and b.size == 256 for a, b in blocks)
    assert np.array_equal(np.concatenate([a for a, _ in blocks])[:1000], x)
    assert np.all(blocks[1][1][44:] == 0.0) and np.all(blocks[3][0][1000 - 768:] == 0.0)
    assert len(list(apio.hop_blocks(x, x, 256, pad_last=False))) == 3


def test_render_signal_against_a_stub_engine():
    """render_signal drives the per-hop loop and concatenates the hop outputs of one rank."""
    class Stub:
        number_of_eigenvectors, hop_size, w_A, w_B = 3, 4, None, None
        def process_input_buffers(self, a, b):
            o = [np.outer(a, [1.0, 2.0]) * (v + 1) for v in range(3)]
            return o, None, [None] * 3, [None] * 3
    x = np.arange(10, dtype=float)
    fa, fb = apio.render_signal(Stub(), x, x, rank=2)
    assert fb is None and fa.shape == (12, 2)
    assert np.array_equal(fa[:10, 0], 2 * x) and np.array_equal(fa[:10, 1], 4 * x) and np.all(fa[10:] == 0)
    with pytest.raises(ValueError):
        apio.render_signal(Stub(), x, x, rank=4)


def test_static_statistics_match_the_vast_m_loops():
    """static_design.static_statistics (vectorised) against the loop restatement of vast.m:42-75."""
    from ap_vast_unofficial_b200.static_design import static_statistics
    from oracle.vast_static_oracle import vast_oracle
    rng = np.random.default_rng(4)
    M, I, L, J, delay, ref = 3, 17, 2, 5, 2, 1
    gB, gD = rng.standard_normal((M, I, L)), rng.standard_normal((M, I, L))
    for N in (40, 12):                       # 12 < I: the reference's fixed horizon truncates the responses
        RB, RD, rB = static_statistics(gB, gD, J, delay, ref, n_samples=N)
        _, RBo, RDo, rBo = vast_oracle(gB, gD, J, delay, ref, 1, 1.0, N=N)
        assert np.max(np.abs(RB - RBo)) < 1e-13 and np.max(np.abs(RD - RDo)) < 1e-13 and np.max(np.abs(rB - rBo)) < 1e-13
