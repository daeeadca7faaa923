"""GPU: the multi-block throughput path (pipelined S1-S4 / S5-S7 overlap, asynchronous D2H ring), the device-side
block-range sharding (halo replay, overlap-add tail, gather), and the reference switches added in round 2
(EXPERIMENTAL_REGULARIZATION off, filters longer than the block, several devices in one process)."""
import ctypes as C

import numpy as np
import pytest

from tests._golden import rel

pytestmark = pytest.mark.gpu


def _engine():
    from ap_vast_unofficial_b200 import apvast
    return apvast


def _case(L=4, M=3, J=16, K=64, Nb=128, H=None, N=192, V=12, nblk=12, seed=31):
    rng = np.random.default_rng(seed)
    dec = np.exp(-np.arange(K) / (K / 4.0)).reshape(-1, 1, 1)
    rA = 1e-3 * rng.standard_normal((K, L, M)) * dec
    rB = 1e-3 * rng.standard_normal((K, L, M)) * dec
    cfg = dict(block_size=Nb, filter_length=J, modeling_delay=4, reference_index_A=1, reference_index_B=2,
               number_of_eigenvectors=V, mu=1.0, statistics_buffer_length=N, perceptual=False)
    if H is not None:
        cfg["hop_size"] = H
    hop = H or Nb // 2
    sA, sB = rng.standard_normal(nblk * hop), rng.standard_normal(nblk * hop)
    return rA, rB, cfg, sA, sB, nblk, hop


@pytest.mark.parametrize("shape", [dict(), dict(L=4, J=64, K=128, Nb=256, N=384, V=16, nblk=8),
                                   dict(L=8, J=128, K=256, Nb=512, N=640, V=32, M=4, nblk=6, eig_mode=3)])
def test_process_blocks_pipelined_is_bit_identical_to_the_per_hop_loop(shape):
    """apv_process_blocks (S1-S4 of hop t+1 on a second stream while S5-S7 of hop t run; outputs through the HBM ring,
    the copy stream and the pinned ring) == a loop of apv_process_block, bit for bit, with and without the overlap."""
    shape = dict(shape)
    eig_mode = shape.pop("eig_mode", 0)
    rA, rB, cfg, sA, sB, nblk, H = _case(**shape)
    V, L = cfg["number_of_eigenvectors"], rA.shape[1]
    np.random.seed(0); e1 = _engine()(rir_A=rA, rir_B=rB, eig_mode=eig_mode, **cfg)
    ref_A, ref_B, ref_At, ref_w = [], [], [], []
    for t in range(nblk):
        oA, oB, oAt, oBt = e1.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        ref_A.append(np.stack(oA)); ref_B.append(np.stack(oB)); ref_At.append(oAt[0].copy())
        ref_w.append(np.stack([e1.w_A[:, :, 0], e1.w_B[:, :, 0]]))
    for pipe, depth in ((True, 4), (True, 3), (True, 2), (True, 1), (1, 2), (1, 1), (False, 1)):
        np.random.seed(0); e2 = _engine()(rir_A=rA, rir_B=rB, eig_mode=eig_mode, **cfg)
        e2.set_pipeline(pipe)
        e2.set_depth(depth)          # joint diagonalisations of two consecutive blocks side by side, or one at a time
        # two calls: the second one starts from the state the first one left (the streams must hand over correctly)
        k = nblk // 2
        a1 = e2.process_blocks(sA[:k * H], sB[:k * H], want_filters=True)
        a2 = e2.process_blocks(sA[k * H:], sB[k * H:], want_filters=True)
        oA = np.concatenate([a1[0], a2[0]]); oB = np.concatenate([a1[1], a2[1]])
        oAt = np.concatenate([a1[2], a2[2]]); w = np.concatenate([a1[4], a2[4]])
        for t in range(nblk):
            assert np.array_equal(oA[t], ref_A[t]), (pipe, depth, t)
            assert np.array_equal(oB[t], ref_B[t]), (pipe, depth, t)
            assert np.array_equal(oAt[t], ref_At[t]), (pipe, depth, t)
            assert np.array_equal(w[t], ref_w[t]), (pipe, depth, t)
        # the handle is left in the per-block layout: attributes and the next per-hop call see the last block
        assert np.array_equal(e2.w_A[:, :, 0], ref_w[-1][0])
        assert np.array_equal(e2.R_A_to_A, e1.R_A_to_A)
        e2.close()
    e1.close()


@pytest.mark.parametrize("hop_div", [2, 4])
def test_range_runner_two_ranks_emulated_on_one_gpu(hop_div):
    """Device path of the block-range sharding (apv_range_run / tail / gather) with two ranks run one after the other
    on one GPU: halo replay of S1-S3, the packed overlap-add tail handed over through the test hooks (the NCCL
    send/recv needs two GPUs: scripts/run_sharded_nccl.py), gather.  hop = Nb/4 needs the longer halo
    ceil(N/H) - 1 + 2 (Nb/H - 1) + ceil((K-1)/H) and a tail that reaches into three blocks."""
    from ap_vast_unofficial_b200 import _capi as capi
    from ap_vast_unofficial_b200.sharded import RangeRunner, tail_blocks, warmup_blocks
    Nb = 128
    rA, rB, cfg, sA, sB, nblk, H = _case(Nb=Nb, H=Nb // hop_div, nblk=20 if hop_div == 2 else 30, seed=33)
    V, L, n = cfg["number_of_eigenvectors"], rA.shape[1], rA.shape[1] * cfg["filter_length"]
    lib = capi.lib()
    np.random.seed(0); ref = _engine()(rir_A=rA, rir_B=rB, **cfg)
    out_ref, w_ref = [], []
    for t in range(nblk):
        oA, oB, _, _ = ref.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        out_ref.append(np.stack([np.stack(oA), np.stack(oB)])); w_ref.append(np.stack([ref.w_A[:, :, 0], ref.w_B[:, :, 0]]))
    wu = warmup_blocks(cfg["statistics_buffer_length"], H, rA.shape[0], Nb)
    assert tail_blocks(Nb, H) == hop_div - 1
    k0 = nblk // 2
    assert k0 >= wu
    np.random.seed(0); e0 = _engine()(rir_A=rA, rir_B=rB, **cfg)
    np.random.seed(0); e1 = _engine()(rir_A=rA, rir_B=rB, **cfg)
    r0 = RangeRunner(e0, 0, 1, None, max_owned=k0)
    r1 = RangeRunner(e1, 0, 1, None, max_owned=nblk - k0)
    assert r1.halo == wu
    r0.run(sA[:k0 * H], sB[:k0 * H], 0, k0)
    r1.run(sA[(k0 - wu) * H:], sB[(k0 - wu) * H:], wu, nblk - k0)
    o0, w0 = np.empty((k0, 2, V, H, L)), np.empty((k0, 2, V, n))
    o1, w1 = np.empty((nblk - k0, 2, V, H, L)), np.empty((nblk - k0, 2, V, n))
    r0.gather([k0], o0, w0)
    r1.gather([nblk - k0], o1, w1)
    # without the tail the first Nb/H - 1 blocks of rank 1 are wrong: the halo exchange is not vacuous
    for k in range(hop_div - 1):
        assert rel(o1[k], out_ref[k0 + k]) > 1e-6
    assert rel(o1[hop_div - 1], out_ref[k0 + hop_div - 1]) <= 1e-12
    tail = np.zeros(2 * V * L * (Nb - H))
    capi.check(lib.apv_range_tail_get(e0._h, capi.ptr(tail)))
    capi.check(lib.apv_range_tail_add(e1._h, capi.ptr(tail)))
    r1.gather([nblk - k0], o1, w1)
    out, w = np.concatenate([o0, o1]), np.concatenate([w0, w1])
    for t in range(nblk):
        assert rel(out[t], out_ref[t]) <= 1e-12, t
        assert rel(w[t], w_ref[t]) <= 1e-12, t
    for e in (ref, e0, e1):
        e.close()


def test_experimental_regularization_off_matches_the_oracle():
    """EXPERIMENTAL_REGULARIZATION = False (apvast.py:25-27): dark += 1e-8 |R_D|_2 I, spectral norm on the device."""
    import importlib
    import oracle.apvast_oracle as ora
    mod = importlib.import_module("ap_vast_unofficial_b200.apvast")   # the module, not the class the package re-exports
    rA, rB, cfg, sA, sB, nblk, H = _case(nblk=7, seed=5)
    mod.EXPERIMENTAL_REGULARIZATION = False
    ora.EXPERIMENTAL_REGULARIZATION = False
    try:
        np.random.seed(0); gpu = mod.apvast(rir_A=rA, rir_B=rB, **cfg)
        np.random.seed(0); o = ora.ApvastOracle(rir_A=rA, rir_B=rB, **cfg)
        for t in range(nblk):
            og = gpu.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
            oo = o.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
            lam = o.lambda_A
            assert np.max(np.abs(gpu.lambda_A - lam[:len(gpu.lambda_A)])) / lam[0] < 1e-10
            gap = np.abs(np.diff(lam[:cfg["number_of_eigenvectors"] + 1])) / lam[0]
            for v in range(cfg["number_of_eigenvectors"]):
                if gap[v] > 1e-9:
                    assert rel(gpu.w_A[v], o.w_A[v]) < 1e-8, (t, v)
            assert rel(np.array(og[0]), np.array(oo[0])) < 1e-8
        # the flag is read at call time, like the reference's jdiag does
        mod.EXPERIMENTAL_REGULARIZATION = True
        ora.EXPERIMENTAL_REGULARIZATION = True
        a, b = sA[:H], sB[:H]
        gpu.process_input_buffers(a, b); o.process_input_buffers(a, b)
        assert rel(gpu.w_A[0], o.w_A[0]) < 1e-8
        gpu.close()
    finally:
        mod.EXPERIMENTAL_REGULARIZATION = True
        ora.EXPERIMENTAL_REGULARIZATION = True


def test_filter_longer_than_the_block_is_cropped_like_rfft():
    """filter_length > block_size: the reference's rfft(w, Nb) keeps the first Nb taps (apvast.py:417-420)."""
    from oracle.apvast_oracle import ApvastOracle
    rA, rB, cfg, sA, sB, nblk, H = _case(L=3, M=2, J=40, K=24, Nb=32, N=96, V=5, nblk=8, seed=12)
    np.random.seed(0); gpu = _engine()(rir_A=rA, rir_B=rB, **cfg)
    np.random.seed(0); o = ApvastOracle(rir_A=rA, rir_B=rB, **cfg)
    for t in range(nblk):
        og = gpu.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        oo = o.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        assert rel(np.array(og[0]), np.array(oo[0])) < 1e-8, t
        assert rel(np.array(og[1]), np.array(oo[1])) < 1e-8, t
    gpu.close()


@pytest.mark.parametrize("shape", [dict(L=3, J=7, M=2, Nb=48, H=16, N=80, V=4, K=20),
                                   dict(L=5, J=9, M=2, Nb=60, H=20, N=90, V=6, K=33)])
def test_render_odd_shapes_against_oracle(shape):
    """The tiled renderer (hop tiles of 64, loudspeaker pairs, double2 stores) on shapes that divide nothing."""
    from oracle.apvast_oracle import ApvastOracle
    rA, rB, cfg, sA, sB, nblk, H = _case(nblk=9, seed=77, **shape)
    np.random.seed(0); gpu = _engine()(rir_A=rA, rir_B=rB, **cfg)
    np.random.seed(0); o = ApvastOracle(rir_A=rA, rir_B=rB, **cfg)
    for t in range(nblk):
        og = gpu.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        oo = o.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        for i in range(4):
            assert rel(np.array(og[i]), np.array(oo[i])) < 1e-8 or np.linalg.norm(np.array(oo[i])) == 0, (t, i)
        assert rel(gpu.output_A_overlap_buffer, o.output_A_overlap_buffer) < 1e-8
    gpu.close()


def test_two_engines_on_two_devices_in_one_process():
    """The handle carries its device: every entry point switches to it and restores the caller's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rA, rB, cfg, sA, sB, nblk, H = _case(L=8, J=128, K=256, Nb=512, N=640, V=32, M=4, nblk=4)   # > 48 KB kernels
    np.random.seed(0); e0 = _engine()(rir_A=rA, rir_B=rB, device=0, **cfg)
    np.random.seed(0); e1 = _engine()(rir_A=rA, rir_B=rB, device=1, **cfg)
    for t in range(nblk):
        a, b = sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H]
        o0 = e0.process_input_buffers(a, b)
        o1 = e1.process_input_buffers(a, b)
        assert np.array_equal(np.array(o0[0]), np.array(o1[0]))
    assert np.array_equal(e0.w_B, e1.w_B)
    e0.close(); e1.close()


@pytest.mark.parametrize("Z", [3])
def test_multizone_per_zone_perceptual_weighting(Z):
    """cfg-5 semantics with perceptual=True: every microphone weighted from its own zone's target (generalisation of
    apvast.py:259-262,318-319); engines exchange weighting curves on the device between S2 and S3."""
    from ap_vast_unofficial_b200.zones import apvast_zones
    from oracle.multizone_oracle import MultiZoneOracle
    rng = np.random.default_rng(50 + Z)
    K, L, M, J, V = 24, 3, 2, 6, 6
    dec = np.exp(-np.arange(K) / 6.0).reshape(-1, 1, 1)
    rirs = [1e-3 * rng.standard_normal((K, L, M)) * dec for _ in range(Z)]
    refs = [z % L for z in range(Z)]
    cfg = dict(block_size=96, filter_length=J, modeling_delay=2, number_of_eigenvectors=V, mu=0.5,
               statistics_buffer_length=144)
    np.random.seed(1)
    gpu = apvast_zones(rirs=rirs, reference_indices=refs, perceptual=True, **cfg)
    ora = MultiZoneOracle(rirs=rirs, reference_indices=refs, seed=1, perceptual=True, **cfg)
    for t in range(10):
        xs = [rng.standard_normal(48) for _ in range(Z)]
        gpu.process_input_buffers(xs)
        ora.process_input_buffers(xs)
    for z in range(Z):
        RB, RD, rB = gpu.statistics(z)
        assert rel(RB, ora.R_B[z]) < 1e-10 and rel(RD, ora.R_D[z]) < 1e-10 and rel(rB[:, 0], ora.r_B[z]) < 1e-10
        w = gpu.w[z][:, :, 0]
        gap = np.abs(np.diff(ora.lam[z][:V + 1])) / ora.lam[z][0]
        for v in range(V):
            if gap[v] > 1e-9:
                assert rel(w[v], ora.w[z][v]) < 1e-8, (z, v)
    gpu.close()


@pytest.mark.parametrize("solver", ["dc", "invit"])
def test_full_spectrum_path_at_cfg2_size(solver, monkeypatch):
    """V = n = 1024 on both zones: the many-vector route (divide and conquer on the tridiagonal matrix -- or, with
    APV_EIG_NO_DC, multisection + lane-per-vector inverse iteration --, aggregated block-reflector GEMM
    back-transformation, GEMM back substitution; BASELINE cfg-4 semantics).  Checks: jdiag.m:33-35 identities,
    the closed form w[V-1] = (R_B + mu (R_D + reg I))^-1 r_B (apVast.m:115-118), the one-launch mu sweep and its
    eigen-basis figures of merit against direct evaluation."""
    from ap_vast_unofficial_b200.workloads import make_workload
    if solver == "invit":
        monkeypatch.setenv("APV_EIG_NO_DC", "1")
    else:
        monkeypatch.delenv("APV_EIG_NO_DC", raising=False)
    wl = make_workload("cfg2", n_blocks=5)
    cfg = dict(wl["cfg"]); n = 1024
    cfg["number_of_eigenvectors"] = n
    np.random.seed(0)
    eng = _engine()(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **cfg)
    H = eng.hop_size
    for t in range(5):
        eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
    mus = np.array([0.01, 1.0, 10.0])
    wA, wB = eng.sweep(mus)
    met = eng.sweep_metrics(mus)
    for (RB, RD, U, lam, r, w, zi) in ((eng.R_A_to_A, eng.R_A_to_B, eng.U_A, eng.lambda_A, eng.r_A[:, 0], wA, 0),
                                       (eng.R_B_to_B, eng.R_B_to_A, eng.U_B, eng.lambda_B, eng.r_B[:, 0], wB, 1)):
        assert np.all(np.diff(lam) <= 0)
        Bm = RD + 1e-7 * np.eye(n)
        import scipy.linalg as sla
        lref = sla.eigh(RB, Bm, eigvals_only=True)[::-1]
        assert np.max(np.abs(lam - lref)) / lref[0] < 1e-11
        assert np.max(np.abs(U.T @ Bm @ U - np.eye(n))) < 1e-8
        assert np.max(np.abs(U.T @ RB @ U - np.diag(lam))) / lam[0] < 1e-8
        for k, mu in enumerate(mus):
            closed = np.linalg.solve(RB + mu * Bm, r)
            assert rel(w[k, -1], closed) < 1e-9, (zi, k)
            for v in (0, 17, 500, n - 1):
                x = w[k, v]
                assert abs(met[k, zi, v, 0] - x @ Bm @ x) <= 1e-8 * abs(x @ Bm @ x)
                assert abs(met[k, zi, v, 1] - x @ RB @ x) <= 1e-8 * abs(x @ RB @ x)
                assert abs(met[k, zi, v, 2] - x @ r) <= 1e-8 * abs(x @ r)
    # the per-block filters (mu of the constructor) are the same prefix sums
    assert rel(eng.w_A[:, :, 0], eng.sweep([eng.mu])[0][0]) < 1e-12
    eng.close()


@pytest.mark.parametrize("env", [dict(APV_SYRK_SLOTS="2"), dict(APV_SYRK_GROUP="4"), dict(APV_SYRK_GROUP="8", APV_SYRK_SLOTS="3")])
def test_syrk_ring_of_partial_tiles_is_bit_identical_whatever_its_size(env, monkeypatch):
    """The statistics SYRK sums the per-microphone partial tiles in the CTA that finishes a tile last, out of a small ring
    of slots (csrc/stats.cu).  The order of the sum is fixed by the code, so R must not depend on the ring size (2 slots:
    nearly every CTA has to wait for its slot), nor on the number of microphones per launch as long as it is a multiple of the
    tree width 4 (later launches accumulate into R), nor on the run: bitwise equal to the default."""
    rA, rB, cfg, sA, sB, nblk, H = _case(L=8, J=64, K=128, Nb=512, N=640, V=16, M=9, nblk=4)   # n = 512: 10 tiles x 4 paths
    apvast = _engine()

    def run():
        np.random.seed(5)
        eng = apvast(rir_A=rA, rir_B=rB, device=0, **cfg)
        for t in range(nblk):
            eng.process_input_buffers(sA[t * H:(t + 1) * H], sB[t * H:(t + 1) * H])
        out = [np.array(getattr(eng, k)) for k in ("R_A_to_A", "R_A_to_B", "R_B_to_A", "R_B_to_B", "w_A", "w_B")]
        eng.close()
        return out

    ref = run()
    again = run()
    for a, b in zip(ref, again):
        assert np.array_equal(a, b)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run()
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert np.allclose(ref[0], ref[0].T, rtol=0, atol=0)          # mirrored lower triangle


def test_bench_line_of_the_gpu_arm_carries_the_contract_keys():
    """`python bench.py` (small workload, so that it takes seconds): ONE JSON line with the driver's keys -- metric / value /
    e2e with host-copy bytes / gpu_launches / clocks -- and the tier's `roofline` and `cpu_baseline` objects."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "small", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "filter_updates_per_sec" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 4
    assert d["value"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"]
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["roofline"]["bound"] in ("hbm", "tensor") and 0 < d["roofline"]["frac"] < 1.5
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
