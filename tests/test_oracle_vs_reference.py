"""CPU: the oracle restatement against the LIVE unmodified reference (only where /root/reference is mounted;
the committed golden vectors cover the GPU box, where it is not)."""
import numpy as np
import pytest

from oracle.ref_import import have_reference, load_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="reference not mounted")


@pytest.mark.parametrize("kw", [dict(), dict(hop_size=24), dict(run_A=False)])
def test_oracle_tracks_live_reference(kw):
    from oracle.apvast_oracle import ApvastOracle
    ref = load_reference()
    rng = np.random.default_rng(17)
    K, L, M = 30, 3, 2
    rA = 1e-3 * rng.standard_normal((K, L, M)); rB = 1e-3 * rng.standard_normal((K, L, M))
    cfg = dict(block_size=96, filter_length=6, modeling_delay=2, reference_index_A=0, reference_index_B=1,
               number_of_eigenvectors=5, mu=0.8, statistics_buffer_length=120, perceptual=False)
    np.random.seed(3); a = ref.apvast(rir_A=rA, rir_B=rB, **cfg, **kw)
    np.random.seed(3); b = ApvastOracle(rir_A=rA, rir_B=rB, **cfg, **kw)
    H = a.hop_size
    for t in range(6):
        xa, xb = rng.standard_normal(H), rng.standard_normal(H)
        oa, ob = a.process_input_buffers(xa, xb), b.process_input_buffers(xa, xb)
        for nm in ("R_B_to_B", "R_B_to_A", "r_B", "w_B"):
            x, y = getattr(a, nm), getattr(b, nm)
            assert np.linalg.norm(x - y) <= 1e-9 * np.linalg.norm(x), (t, nm)
        for x, y in zip(oa, ob):
            if x is not None:
                assert np.linalg.norm(np.array(x) - np.array(y)) <= 1e-9 * max(np.linalg.norm(np.array(x)), 1e-300)
