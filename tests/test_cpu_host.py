"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ap_vast_unofficial_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "apvast_b200.h")).read()
    declared = set(re.findall(r"\b(apv_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"apv_set_weights"}          # mentioned in a comment only
    assert len(declared) >= 25
    lib = _capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    assert b"sm_100a" in lib.apv_version()


def test_config_struct_layout_matches_header():
    from ap_vast_unofficial_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "apvast_b200.h")).read()
    body = hdr[hdr.index("typedef struct apv_config {"):hdr.index("} apv_config;")]
    names = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        m = re.match(r"(int32_t|double)\s+([^;]+);", line)
        if m:
            names += [x.strip() for x in m.group(2).split(",")]
    assert names == [f[0] for f in _capi.Config._fields_]
    assert C.sizeof(_capi.Config) == 18 * 4 + 3 * 8 + 4 * 4 + 2 * 8 + 2 * 4


def test_library_links_no_vendor_math_libraries():
    import subprocess
    from ap_vast_unofficial_b200 import _capi
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout.lower()
    for bad in ("cublas", "cusolver", "cufft", "cusparse", "nccl"):
        assert bad not in out


def test_no_gpu_fails_loudly_not_silently():
    """Without a CUDA device the engine must raise, never fall back to a CPU path."""
    from ap_vast_unofficial_b200 import apvast
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    rng = np.random.default_rng(0)
    r = 1e-3 * rng.standard_normal((8, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        apvast(64, r, r, 4, 1, 0, 0, 2, 1.0, 32, perceptual=False)


def test_reference_validation_errors_raised_on_host():
    from ap_vast_unofficial_b200 import apvast
    r = np.zeros((8, 2, 2))
    with pytest.raises(RuntimeError, match="block size must be modulo 2"):       # apvast.py:86-87
        apvast(63, r, r, 4, 1, 0, 0, 2, 1.0, 32, perceptual=False)
    with pytest.raises(RuntimeError, match="rirs of unequal size"):              # apvast.py:89-90
        apvast(64, r, np.zeros((8, 2, 3)), 4, 1, 0, 0, 2, 1.0, 32, perceptual=False)


def test_product_package_never_imports_the_oracle():
    pk = os.path.join(ROOT, "ap_vast_unofficial_b200")
    for fn in os.listdir(pk):
        if fn.endswith(".py"):
            src = open(os.path.join(pk, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_masking_tables_match_matlab_restatement():
    from ap_vast_unofficial_b200.perceptual import MaskingModel
    from oracle.perceptual_oracle import PerceptualModelOracle
    for nb in (96, 1600, 2048):
        a, b = MaskingModel(nb, 48000), PerceptualModelOracle(nb, 48000)
        assert a.n_channels == b.n_channels == 44                    # C = 44 at 48 kHz
        assert abs(a.Cs - b.Cs) < 1e-9 * b.Cs and abs(a.Ca - b.Ca) < 1e-9 * b.Ca
        x = 1e-3 * np.random.default_rng(nb).standard_normal(nb)
        assert np.max(np.abs(a.gain(x) - b.gain(x))) < 1e-12 * np.max(b.gain(x))


def test_perceptual_model_quiet_threshold_scenario():
    """testPerceptualModel.m:21-34 scenario: with (almost) no masker the masking curve follows the ISO 226
    threshold in quiet up to the model's calibration offset, i.e. its shape is the threshold curve."""
    from oracle.perceptual_oracle import PerceptualModelOracle, threshold_of_hearing_db
    nb, fs = 2048, 48000
    m = PerceptualModelOracle(nb, fs)
    sq = m.squared_weighting_curve(np.zeros(nb))
    f = np.arange(nb // 2 + 1) * fs / nb
    band = (f > 200) & (f < 8000)
    mask_db = 10 * np.log10(1.0 / sq[band])
    thr_db = threshold_of_hearing_db(f[band])
    d = mask_db - thr_db
    assert np.max(d) - np.min(d) < 12.0          # same shape within the gammatone smoothing


def test_workloads_are_deterministic_and_named_shapes():
    from ap_vast_unofficial_b200.workloads import make_workload
    a, b = make_workload("cfg2", n_blocks=2), make_workload("cfg2", n_blocks=2)
    assert np.array_equal(a["rir_A"], b["rir_A"]) and np.array_equal(a["signal_B"], b["signal_B"])
    assert a["shapes"]["n"] == 1024 and make_workload("cfg3", n_blocks=1)["shapes"]["n"] == 4096
    assert abs(np.sqrt(np.mean(a["signal_A"] ** 2)) - 1.0) < 1e-12


def test_fft_plan_covers_reference_block_sizes():
    # 1600 = 2^6 5^2 (cfg-1) needs radix 5; the planner must factor any even block size
    def plan(n):
        out, m = [], n
        while m % 4 == 0:
            out.append(4); m //= 4
        p = 2
        while m > 1:
            if m % p == 0:
                out.append(p); m //= p
            else:
                p += 1 if p == 2 else 2
                if p * p > m:
                    p = m
        return out
    for n in (1600, 2048, 1020, 210, 2 * 509):
        assert int(np.prod(plan(n))) == n


def test_divide_and_conquer_prototype_against_lapack():
    """scripts/proto_dc.py is the NumPy restatement of csrc/dc.cu (same merges, deflation scan, secular solver and
    Loewner recomputation): eigenvalues, orthogonality and residual against scipy on a random matrix, on one with
    tiny couplings (everything deflates) and on a pencil-like spectrum with a long tail of tiny eigenvalues."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import proto_dc
    from scipy.linalg import eigh_tridiagonal, hessenberg
    rng = np.random.default_rng(3)
    cases = [(rng.standard_normal(96), rng.standard_normal(95)), (rng.standard_normal(128), 1e-12 * rng.standard_normal(127))]
    lamt = np.concatenate([np.logspace(1, -3, 20), 1e-9 * rng.random(108)])
    Qr, _ = np.linalg.qr(rng.standard_normal((128, 128)))
    Hh = hessenberg(((Qr * lamt) @ Qr.T + ((Qr * lamt) @ Qr.T).T) / 2)
    cases.append((np.diag(Hh).copy(), np.diag(Hh, 1).copy()))
    for d, e in cases:
        lam, Q = proto_dc.dc_eigh(d, e)
        ref = eigh_tridiagonal(d, e, eigvals_only=True)
        n = d.size
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        nrm = np.max(np.abs(ref))
        assert np.max(np.abs(lam - ref)) <= 1e-13 * nrm
        assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13
        assert np.max(np.abs(T @ Q - Q * lam[None, :])) <= 1e-13 * nrm


def _run_bench(args, env_extra=None, launcher=None):
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = (launcher or [sys.executable]) + [os.path.join(root, "bench.py")] + args
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    return [json.loads(l) for l in lines]


def test_bench_reference_arm_prints_one_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line, the arm's own metric /
    unit / config, `impl`, a `cpu_baseline` describing the run and an `e2e` with zero copy bytes."""
    lines = _run_bench(["--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "3"])
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "filter_updates_per_sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("small")


def test_bench_reference_arm_under_torchrun_rank0_alone_prints_and_keeps_its_blas_threads():
    """Under torch.distributed.run (N > 1) rank 0 alone runs the arm and the other ranks exit 0 without work; the launcher
    exports OMP_NUM_THREADS=1, which the arm must not inherit (that made it time out at N > 1 in round 1)."""
    import os
    import sys
    launcher = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", "29541"]
    lines = _run_bench(["--impl", "reference", "--gpus", "2", "--workload", "small", "--steps", "2", "--warmup", "3"],
                       launcher=launcher)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    if (os.cpu_count() or 1) > 1:
        assert lines[0]["cpu_baseline"]["cores"] > 1
