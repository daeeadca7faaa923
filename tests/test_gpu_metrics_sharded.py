"""GPU: (1) acoustic contrast / normalised signal distortion of the rendered feeds within 0.01 dB of the oracle
(north_star criterion); (2) block-range sharding with the CUDA engine, two ranks emulated on one GPU."""
import numpy as np
import pytest

from tests._golden import rel

pytestmark = pytest.mark.gpu


def _case(nblk=14, seed=31):
    rng = np.random.default_rng(seed)
    K, L, M = 64, 4, 3
    dec = np.exp(-np.arange(K) / 16.0).reshape(-1, 1, 1)
    rA = 1e-3 * rng.standard_normal((K, L, M)) * dec
    rB = 1e-3 * rng.standard_normal((K, L, M)) * dec
    cfg = dict(block_size=128, filter_length=16, modeling_delay=4, reference_index_A=1, reference_index_B=2,
               number_of_eigenvectors=12, mu=1.0, statistics_buffer_length=192, perceptual=False)
    sA, sB = rng.standard_normal(nblk * 64), rng.standard_normal(nblk * 64)
    return rA, rB, cfg, sA, sB, nblk


def test_contrast_and_distortion_within_001_db():
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200.metrics import evaluate_zone
    from oracle.apvast_oracle import ApvastOracle
    rA, rB, cfg, sA, sB, nblk = _case()
    feeds = {}
    for name, cls in (("gpu", apvast), ("ora", ApvastOracle)):
        np.random.seed(0)
        eng = cls(rir_A=rA, rir_B=rB, **cfg)
        oa, ob = [], []
        for t in range(nblk):
            A, B, _, _ = eng.process_input_buffers(sA[t * 64:(t + 1) * 64], sB[t * 64:(t + 1) * 64])
            oa.append(np.stack(A)); ob.append(np.stack(B))
        feeds[name] = (np.concatenate(oa, axis=1), np.concatenate(ob, axis=1))     # (V, T, L)
    for v in (0, 5, 11):
        for z, (rb, rd, sig, ref) in enumerate(((rA, rB, sA, 1), (rB, rA, sB, 2))):
            ac_g, nsd_g = evaluate_zone(feeds["gpu"][z][v][5 * 64:], rb, rd, sig[5 * 64:], ref, 4)
            ac_o, nsd_o = evaluate_zone(feeds["ora"][z][v][5 * 64:], rb, rd, sig[5 * 64:], ref, 4)
            assert abs(ac_g - ac_o) < 0.01, (v, z, ac_g, ac_o)
            assert abs(nsd_g - nsd_o) < 0.01, (v, z, nsd_g, nsd_o)


def test_cfg1_contrast_and_distortion_against_reference_feeds():
    """North-star criterion on the reference's own fixture: the make_python_test.m case on Python/rirs.mat, 16 hops.
    AC / NSD of the GPU engine's feeds at the control microphones of rirs.mat within 0.01 dB of the AC / NSD of the feeds
    the UNMODIFIED reference rendered (tests/golden/cfg1_feeds.npz, oracle/make_golden_cfg1_feeds.py), evaluated both on
    the device (apv_eval_zone) and with the host definitions."""
    import os
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200.metrics import evaluate_zone
    from tests._golden import GOLDEN, load_case
    g1, _, _ = load_case("cfg1")
    z = np.load(os.path.join(GOLDEN, "cfg1_feeds.npz"))
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg_")}
    rA, rB = g1["rir_A"], g1["rir_B"]
    nblk, skip, ranks = int(z["nblk"]), int(z["skip"]), list(z["ranks"])
    np.random.seed(int(z["seed"]))
    eng = apvast(rir_A=rA, rir_B=rB, perceptual=False, **cfg)
    H = eng.hop_size
    sA, sB = z["input_A"], z["input_B"]
    fa, fb = [], []
    for t in range(nblk):
        oA, oB, _, _ = eng.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        fa.append(np.stack([oA[v] for v in ranks])); fb.append(np.stack([oB[v] for v in ranks]))
    fa, fb = np.concatenate(fa, axis=1), np.concatenate(fb, axis=1)
    worst = 0.0
    for i in range(len(ranks)):
        assert rel(fa[i], z["feeds_A"][i]) < 1e-8 and rel(fb[i], z["feeds_B"][i]) < 1e-8
        for zi, (f, rb, rd, sig, zone) in enumerate(((fa[i], rA, rB, sA, "A"), (fb[i], rB, rA, sB, "B"))):
            ac_ref, nsd_ref = z["metrics"][i, zi]
            ac_h, nsd_h = evaluate_zone(f[skip * H:], rb, rd, sig[skip * H:], cfg["reference_index_A"], cfg["modeling_delay"])
            ac_d, _, nsd_d = eng.evaluate(f[skip * H:], sig[skip * H:], zone)
            worst = max(worst, abs(ac_h - ac_ref), abs(nsd_h - nsd_ref), abs(ac_d - ac_ref), abs(nsd_d - nsd_ref))
            assert abs(ac_h - ac_ref) < 0.01 and abs(nsd_h - nsd_ref) < 0.01, (i, zone, ac_h, ac_ref, nsd_h, nsd_ref)
            assert abs(ac_d - ac_ref) < 0.01 and abs(nsd_d - nsd_ref) < 0.01, (i, zone, ac_d, ac_ref, nsd_d, nsd_ref)
    print("cfg-1 AC/NSD vs reference feeds: worst |delta| = %.2e dB" % worst)
    eng.close()


def test_device_metrics_match_host_definitions():
    """apv_eval_zone (pressure + energies on the device) against the NumPy definitions of metrics.py."""
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200.metrics import evaluate_zone
    rA, rB, cfg, sA, sB, nblk = _case()
    np.random.seed(0)
    eng = apvast(rir_A=rA, rir_B=rB, **cfg)
    oa, ob = [], []
    for t in range(nblk):
        A, B, _, _ = eng.process_input_buffers(sA[t * 64:(t + 1) * 64], sB[t * 64:(t + 1) * 64])
        oa.append(A[-1]); ob.append(B[-1])
    fa, fb = np.concatenate(oa, axis=0), np.concatenate(ob, axis=0)
    for zone, f, rb, rd, sig, ref in (("A", fa, rA, rB, sA, 1), ("B", fb, rB, rA, sB, 2)):
        ac_h, nsd_h = evaluate_zone(f, rb, rd, sig, ref, 4)
        ac_d, nmse_d, nsd_d = eng.evaluate(f, sig, zone)
        assert abs(ac_d - ac_h) < 1e-9 and abs(nsd_d - nsd_h) < 1e-9, (zone, ac_d, ac_h, nsd_d, nsd_h)


class _FakeDist:
    """Two ranks run one after the other on one GPU; point-to-point messages go through a dict."""
    box = {}

    def __init__(self, rank):
        self.rank = rank

    def get_backend(self):
        return "gloo"

    def get_world_size(self):
        return 2

    class _Req:
        def wait(self):
            return None

    def isend(self, t, dst):
        _FakeDist.box[(self.rank, dst)] = t.clone()
        return self._Req()

    def irecv(self, t, src):
        t.copy_(_FakeDist.box[(src, self.rank)])
        return self._Req()

    def all_gather_object(self, out, obj):
        for i in range(len(out)):
            out[i] = obj if i == self.rank else []


def test_sharded_two_ranks_emulated_matches_single_stream():
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200.sharded import process_signal_sharded
    rA, rB, cfg, sA, sB, nblk = _case(nblk=16, seed=33)
    make = lambda: apvast(rir_A=rA, rir_B=rB, **cfg)
    ref = process_signal_sharded(make, sA, sB, seed=0)
    r0 = process_signal_sharded(make, sA, sB, rank=0, world=2, dist=_FakeDist(0), seed=0, gather=False)
    r1 = process_signal_sharded(make, sA, sB, rank=1, world=2, dist=_FakeDist(1), seed=0, gather=False)
    assert r0["blocks"] == (0, 8) and r1["blocks"] == (8, 16)
    outs = r0["out_A"] + r1["out_A"]
    ws = r0["w_A"] + r1["w_A"]
    for t in range(16):
        assert np.linalg.norm(outs[t] - ref["out_A"][t]) <= 1e-8 * np.linalg.norm(ref["out_A"][t]), t
        for v in range(ws[t].shape[0]):
            assert np.linalg.norm(ws[t][v] - ref["w_A"][t][v]) <= 1e-8 * np.linalg.norm(ref["w_A"][t][v]), (t, v)


def test_render_signal_equals_the_per_hop_loop():
    """io.render_signal = the driver loop of make_python_test.m:44-54 on a whole signal."""
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200 import io as apio
    rng = np.random.default_rng(21)
    K, L, M = 32, 3, 2
    rA = 1e-3 * rng.standard_normal((K, L, M)); rB = 1e-3 * rng.standard_normal((K, L, M))
    cfg = dict(block_size=64, filter_length=8, modeling_delay=2, reference_index_A=0, reference_index_B=1,
               number_of_eigenvectors=5, mu=1.0, statistics_buffer_length=96, perceptual=False)
    x, y = rng.standard_normal(32 * 7 + 5), rng.standard_normal(32 * 7)
    np.random.seed(2); e1 = apvast(rir_A=rA, rir_B=rB, **cfg)
    fa, fb, wA, wB = apio.render_signal(e1, x, y, rank=3, collect_filters=True)
    np.random.seed(2); e2 = apvast(rir_A=rA, rir_B=rB, **cfg)
    ref = []
    for a, b in apio.hop_blocks(x, y, 32):
        oA, oB, _, _ = e2.process_input_buffers(a, b)
        ref.append(np.array(oA[2]))
    assert fa.shape == (8 * 32, L) and fb.shape == (8 * 32, L) and wA.shape == (8, L * 8)
    assert np.array_equal(fa, np.concatenate(ref, axis=0))
    assert np.array_equal(wA[-1], np.array(e2.w_A[2]).reshape(-1))


@pytest.mark.parametrize("V", [1, 7, 24])
def test_static_vast_design_against_the_vast_m_restatement(V):
    """static_design.vast_static (host statistics + apv_jdiag on the GPU) against oracle/vast_static_oracle.py."""
    from ap_vast_unofficial_b200.static_design import vast_static
    from oracle.vast_static_oracle import vast_oracle
    rng = np.random.default_rng(11)
    M, I, L, J = 4, 40, 3, 8
    dec = np.exp(-np.arange(I) / 10.0)[None, :, None]
    gB, gD = rng.standard_normal((M, I, L)) * dec, rng.standard_normal((M, I, L)) * dec
    w = vast_static(gB, gD, J, 3, 1, V, 0.8)
    wo, _, _, _ = vast_oracle(gB, gD, J, 3, 1, V, 0.8)
    assert w.shape == (J, L)
    assert np.linalg.norm(w - wo) / np.linalg.norm(wo) < 1e-8


@pytest.mark.parametrize("Z", [2, 3, 4])
def test_multizone_engine_against_the_oracle_generalisation(Z):
    """zones.apvast_zones (one two-zone engine per bright zone, dark zone = union of the other zones' microphones)
    against oracle/multizone_oracle.py (pairwise two-zone oracles + explicit sums): BASELINE cfg-5 semantics."""
    from ap_vast_unofficial_b200.zones import apvast_zones
    from oracle.multizone_oracle import MultiZoneOracle
    rng = np.random.default_rng(30 + Z)
    K, L, M, J, V = 24, 3, 2, 6, 6
    dec = np.exp(-np.arange(K) / 6.0).reshape(-1, 1, 1)
    rirs = [1e-3 * rng.standard_normal((K, L, M)) * dec for _ in range(Z)]
    refs = [z % L for z in range(Z)]
    cfg = dict(block_size=32, filter_length=J, modeling_delay=2, number_of_eigenvectors=V, mu=0.5,
               statistics_buffer_length=48)
    np.random.seed(1)
    gpu = apvast_zones(rirs=rirs, reference_indices=refs, **cfg)
    ora = MultiZoneOracle(rirs=rirs, reference_indices=refs, seed=1, **cfg)
    for t in range(10):                       # the random start buffers are flushed after (Nb + N) / H + 2 hops
        xs = [rng.standard_normal(16) for _ in range(Z)]
        outs = gpu.process_input_buffers(xs)
        ora.process_input_buffers(xs)
    assert len(outs) == Z and len(outs[0]) == V and outs[0][0].shape == (16, L)
    for z in range(Z):
        RB, RD, rB = gpu.statistics(z)
        assert rel(RB, ora.R_B[z]) < 1e-11 and rel(RD, ora.R_D[z]) < 1e-11 and rel(rB[:, 0], ora.r_B[z]) < 1e-11
        lam = gpu.eigenvalues[z]
        assert np.max(np.abs(lam - ora.lam[z][:V])) / ora.lam[z][0] < 1e-10
        w = gpu.w[z][:, :, 0]
        gap = np.abs(np.diff(ora.lam[z][:V + 1])) / ora.lam[z][0]
        for v in range(V):
            if gap[v] > 1e-9:
                assert rel(w[v], ora.w[z][v]) < 1e-8, (z, v)
    gpu.close()
