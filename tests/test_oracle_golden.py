"""CPU: the NumPy restatement (oracle/) against the golden vectors the unmodified reference produced."""
import numpy as np
import pytest

from oracle.apvast_oracle import ApvastOracle, jdiag, toeplitz_rows
from tests._golden import GOLDEN, compare_state, rel, replay

SMALL = ["tiny", "tiny_hop", "tiny_runA", "tiny_full", "mid"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_golden(name):
    eng, g, res = replay(ApvastOracle, name)
    assert res, "no compared blocks"
    for t, e in res.items():
        for k, v in e.items():
            tol = 1e-8 if k.startswith("w_") or k.startswith("out_") else 1e-10
            assert v <= tol, (name, t, k, v)
    for a, v in compare_state(eng, g).items():
        assert v <= 1e-9, (name, a, v)


def test_oracle_matches_reference_golden_perceptual_injected():
    # gain model injected at the libdetectability boundary in both (parity otherwise unpinned)
    eng, g, res = replay(ApvastOracle, "tiny_perc")
    for t, e in res.items():
        for k, v in e.items():
            assert v <= 1e-8, (t, k, v)
    st = compare_state(eng, g)
    assert st["weighting_spectra_A"] <= 1e-12 and st["weighting_spectra_B"] <= 1e-12


def test_oracle_matches_reference_golden_cfg1():
    eng, g, res = replay(ApvastOracle, "cfg1")
    assert sorted(res) == [1, 5, 9]
    for t, e in res.items():
        for k, v in e.items():
            tol = 1e-8 if k.startswith("w_") or k.startswith("out_") else 1e-10
            assert v <= tol, (t, k, v)


def test_toeplitz_quirk_matches_scipy():
    import scipy.linalg as sla
    rng = np.random.default_rng(0)
    s = rng.standard_normal(40)
    J = 7
    want = sla.toeplitz(np.flipud(s[0:J]), s[J:])      # exactly the reference expression (apvast.py:336-338)
    assert np.array_equal(toeplitz_rows(s, J), want)


def test_jdiag_identities():
    # jdiag.m:33-35: U'AU = D, U'BU = I
    rng = np.random.default_rng(3)
    n = 40
    X = rng.standard_normal((n, 3 * n)); A = X @ X.T
    Y = rng.standard_normal((n, 3 * n)); B = Y @ Y.T
    U, D = jdiag(A, B)
    assert np.allclose(U.T @ (B + 1e-7 * np.eye(n)) @ U, np.eye(n), atol=1e-9)
    assert np.allclose(U.T @ A @ U, D, atol=1e-8 * np.abs(D).max())
    assert np.all(np.diff(np.diag(D)) <= 0)


def test_full_rank_closed_form():
    # apVast.m:115-118 / vast.m:92: V = n  =>  w = (R_B + mu (R_D + reg I))^-1 r_B
    eng, g, res = replay(ApvastOracle, "tiny_full")
    n = eng.R_A_to_A.shape[0]
    w = np.linalg.solve(eng.R_A_to_A + eng.mu * (eng.R_A_to_B + 1e-7 * np.eye(n)), eng.r_A)
    assert np.linalg.norm(w - eng.w_A[-1]) / np.linalg.norm(w) < 1e-8


def test_matlab_flavour_closed_form():
    """flavour='matlab' (apVast.m differences): with V = n the last filter is the pressure-matching
    solution on the LOADED statistics, w = (R_B' + mu R_D')^-1 r_B (apVast.m:115-118, 552-569); no parity pin exists
    for this flavour (no MATLAB/Octave here), so it is checked through identities only."""
    rng = np.random.default_rng(8)
    K, L, M = 20, 3, 2
    rA = 1e-3 * rng.standard_normal((K, L, M)); rB = 1e-3 * rng.standard_normal((K, L, M))
    eng = ApvastOracle(64, rA, rB, 6, 2, 0, 2, 18, 0.7, 90, perceptual=False, flavour="matlab")
    assert np.all(eng.loudspeaker_response_A_to_A_buffer == 0)           # zero start (apVast.m:175-180)
    for t in range(5):
        eng.process_input_buffers(rng.standard_normal(32), rng.standard_normal(32))
    n = 18
    w = np.linalg.solve(eng.R_B_to_B + 0.7 * eng.R_B_to_A, eng.r_B)
    assert np.linalg.norm(w - eng.w_B[-1]) / np.linalg.norm(w) < 1e-8
    assert eng._data_matrix(eng.loudspeaker_weighted_response_A_to_A_buffer, 0).shape == (n, 90 - 6 + 1)
    # target of zone B sits on loudspeaker reference_index_B
    ft = eng.filter_spectra_B_t[0]
    assert np.allclose(np.abs(ft[:, 2]), 1.0) and np.allclose(ft[:, 0], 0.0)


def test_multizone_oracle_reduces_to_the_two_zone_oracle():
    """oracle/multizone_oracle.py (SURVEY 8d cfg-5 generalisation) at Z = 2 is the pinned two-zone oracle."""
    from oracle.apvast_oracle import ApvastOracle
    from oracle.multizone_oracle import MultiZoneOracle
    rng = np.random.default_rng(5)
    K, L, M, J, V = 24, 2, 2, 6, 5
    rirs = [1e-3 * rng.standard_normal((K, L, M)) for _ in range(2)]
    cfg = dict(block_size=32, filter_length=J, modeling_delay=2, number_of_eigenvectors=V, mu=0.5,
               statistics_buffer_length=48)
    mz = MultiZoneOracle(rirs=rirs, reference_indices=[0, 1], seed=3, **cfg)
    np.random.seed(3)
    two = ApvastOracle(rir_A=rirs[0], rir_B=rirs[1], reference_index_A=0, reference_index_B=1, perceptual=False, **cfg)
    for t in range(6):
        a, b = rng.standard_normal(16), rng.standard_normal(16)
        mz.process_input_buffers([a, b])
        two.process_input_buffers(a, b)
    for v in range(V):
        wa, wb = np.array(two.w_A)[v, :, 0], np.array(two.w_B)[v, :, 0]
        assert np.linalg.norm(mz.w[0][v] - wa) / np.linalg.norm(wa) < 1e-9
        assert np.linalg.norm(mz.w[1][v] - wb) / np.linalg.norm(wb) < 1e-9


def test_oracle_against_cfg2_reference_golden():
    """The oracle at BASELINE cfg-2 size (n = 1024) against the UNMODIFIED reference (tests/golden/cfg2_reference.npz,
    oracle/make_golden_cfg3.py --workload cfg2): three hops, all 64 ranks."""
    import os
    from ap_vast_unofficial_b200.workloads import make_workload
    from oracle.apvast_oracle import ApvastOracle
    z = np.load(os.path.join(GOLDEN, "cfg2_reference.npz"))
    nblk = 3
    wl = make_workload("cfg2", n_blocks=int(z["nblk"]))      # (the programme signals are normalised over their whole length)
    np.random.seed(int(z["seed"]))
    eng = ApvastOracle(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **wl["cfg"])
    H, V = eng.hop_size, eng.number_of_eigenvectors
    for t in range(nblk):
        outs = eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
        assert rel(np.diag(eng.R_A_to_B), z[f"R_A_to_B_diag_{t}"]) < 1e-13
        lam_ref = z[f"lambda_A_{t}"]
        gap = np.abs(np.diff(lam_ref)) / lam_ref[0]
        for v in range(V):
            if gap[v] > 1e-9:
                assert rel(eng.w_A[v, :, 0], z[f"w_A_{t}"][v]) < 1e-8, (t, v)
        got, want = np.stack([outs[1][0], outs[1][V - 1]]), z[f"out_B_{t}"]
        # (hop 0 renders 1e-18: the first input block is still almost all zeros)
        assert np.linalg.norm(got - want) <= 1e-9 * max(np.linalg.norm(want), 1e-6), t
