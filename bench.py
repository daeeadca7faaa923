#!/usr/bin/env python
"""Benchmark of the AP-VAST per-block hot path (BASELINE.json metric: filter updates/sec & real-time factor).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2] [--impl ours|reference]

A *step* is one filter update (one ``process_input_buffers`` call = S1..S7 over one hop of both programme
signals) per rank.  The workload is BASELINE.json ``configs[2]`` (L=16, J=256, n=4096 -- the configuration the
north-star target is quoted on; it fits one GPU); ``--workload cfg2`` selects ``configs[1]``.  At N > 1 the
60-s signal is block-range sharded: every rank owns a contiguous range and processes its own blocks with no
collective on the per-block path (weak scaling: per-rank work is fixed).

Printed JSON line (rank 0):
  value     whole-job filter updates/s with the step inputs already resident in HBM (device-timed, CUDA
            events on the engine's stream, max over ranks)
  e2e       the same metric through the public drop-in call ``apvast.process_input_buffers`` with HOST
            buffers: pinned H2D of the hop and D2H of the rendered outputs inside the timed region
  roofline  the dominant kernel.  At cfg3 (n >= 2048: two-stage tridiagonalisation) that is the FP64 tensor-core
            statistics SYRK: algorithmic flops per block / its launch duration against the DMMA peak measured
            live in this process (MEASURED_PEAKS.json has no FP64 figure).  At cfg2 (one-stage tridiagonalisation)
            it is td_panel_kernel against MEASURED_PEAKS.json hbm_gbs.
  roofline_tridiag / roofline_stats   the other of the two
  cpu_baseline     the oracle port of the reference timed on this box's host cores on a bounded sample

``--impl reference`` times the reference algorithm's CPU implementation (the oracle port: /root/reference is
not on the GPU box and the reference is pure NumPy/SciPy) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 48000.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sample(wl, n_timed, n_state_warm, mics_sample=2, verbose=False):
    """Time the oracle port of the reference on a bounded sample of the workload.

    Per timed step: S1-S3, S6, S7 in full; S4 on `mics_sample` of the M microphones (cost is exactly linear in
    the microphones: apvast.py:332-364 loops over m) and S5 on ONE of the two zones (the two jdiag calls,
    apvast.py:380-382, are identical in cost).  Step time = S123 + S4_sample * M/mics_sample + 2 * S5_sample + S67.
    Returns (seconds per block list, per-stage split of the last step)."""
    import scipy.linalg as sla
    from oracle import apvast_oracle as ora
    cfg = wl["cfg"]
    np.random.seed(0)
    eng = ora.ApvastOracle(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, **cfg)
    H = eng.hop_size
    M = eng.number_of_mics
    ms = min(mics_sample, M)
    J, V, mu = eng.filter_length, eng.number_of_eigenvectors, eng.mu
    t_blk = 0
    _w = np.random.default_rng(0).standard_normal((512, 512))
    for _ in range(3):                      # spin up the BLAS thread pool before anything is timed
        _ = _w @ _w.T
    for _ in range(n_state_warm):           # fill the statistics buffers (state only, untimed)
        eng.advance_state(wl["signal_A"][t_blk * H:(t_blk + 1) * H], wl["signal_B"][t_blk * H:(t_blk + 1) * H])
        t_blk += 1
    times, split = [], {}
    for _ in range(n_timed):
        a = wl["signal_A"][t_blk * H:(t_blk + 1) * H]; b = wl["signal_B"][t_blk * H:(t_blk + 1) * H]
        t_blk += 1
        t0 = time.perf_counter()
        eng.update_loudspeaker_response_buffers(a, b)
        eng.update_weighted_target_signals()
        eng.update_weighted_loudspeaker_response()
        t1 = time.perf_counter()
        # S4 sample: zone-A bright/dark + zone-B bright/dark for `ms` microphones
        n = J * eng.number_of_srcs
        R = [np.zeros((n, n)) for _ in range(4)]
        r = [np.zeros((n, 1)) for _ in range(2)]
        for m in range(ms):
            Y = eng._data_matrix(eng.loudspeaker_weighted_response_A_to_A_buffer, m); R[0] += Y @ Y.T
            r[0] += Y @ eng.loudspeaker_weighted_target_response_A_to_A_buffer[J:, m].reshape(-1, 1)
            Y = eng._data_matrix(eng.loudspeaker_weighted_response_A_to_B_buffer, m); R[1] += Y @ Y.T
            Y = eng._data_matrix(eng.loudspeaker_weighted_response_B_to_B_buffer, m); R[3] += Y @ Y.T
            r[1] += Y @ eng.loudspeaker_weighted_target_response_B_to_B_buffer[J:, m].reshape(-1, 1)
            Y = eng._data_matrix(eng.loudspeaker_weighted_response_B_to_A_buffer, m); R[2] += Y @ Y.T
        t2 = time.perf_counter()
        # S5 sample: one zone (regularised a little more because the sampled R_D has fewer microphones)
        U, D = ora.jdiag(R[0], R[1] + 1e-9 * np.trace(R[1]) / n * np.eye(n))
        t3 = time.perf_counter()
        lam = np.diag(D)
        c = U[:, :V].T @ r[0].reshape(-1)
        w = np.cumsum((c / (lam[:V] + mu))[None, :] * U[:, :V], axis=1).T.reshape(V, n, 1)
        eng.w_A = w; eng.w_B = w
        L, Nb = eng.number_of_srcs, eng.block_size
        eng.filter_spectra_A = [np.fft.rfft(w[v, :, 0].reshape(L, J).T, Nb, axis=0) for v in range(V)]
        eng.filter_spectra_B = eng.filter_spectra_A
        ft = np.zeros(n); ft[J * eng.reference_index_A + eng.modeling_delay] = 1.0
        ftf = np.fft.rfft(ft.reshape(L, J).T, Nb, axis=0)
        eng.filter_spectra_A_t = [ftf] * V; eng.filter_spectra_B_t = [ftf] * V
        eng.update_input_blocks(a, b)
        eng.compute_output_buffers()
        t4 = time.perf_counter()
        s123, s4, s5, s67 = t1 - t0, (t2 - t1) * (M / ms), 2.0 * (t3 - t2), t4 - t3
        times.append(s123 + s4 + s5 + s67)
        split = {"S1S2S3": s123, "S4_scaled": s4, "S5_scaled": s5, "S6S7": s67, "sample_wall_s": t4 - t0}
        if verbose:
            print("cpu sample", split, file=sys.stderr)
    return times, split, f"per step: S1-S3,S6,S7 full; S4 on {ms}/{M} mics x{M / ms:g}; S5 (jdiag, n={n}) on 1/2 zones x2"


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from ap_vast_unofficial_b200.workloads import make_workload
    wl = make_workload(args.workload, n_blocks=args.steps + args.warmup + 4)
    sh = wl["shapes"]
    t0 = time.perf_counter()
    # warm-up steps only advance the state (there is nothing to warm on the CPU besides filling the buffers)
    times, split, sample = cpu_reference_sample(wl, args.steps, max(args.warmup, 3))
    sec = float(np.mean(times))
    ups = 1.0 / sec
    line = {
        "impl": "reference", "metric": "filter_updates_per_sec", "value": ups, "unit": "updates/s",
        "rtf": ups * sh["H"] / FS, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: synthetic 2-zone L={sh['L']} M={sh['M']} J={sh['J']} n={sh['n']} "
                               f"K={sh['K']} Nb={sh['Nb']} H={sh['H']} N={sh['N']} V={sh['V']} fs=48000"},
        "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": blas_threads(), "kind": "port",
                         "sample": sample, "split_s": split, "host_cpus": os.cpu_count()},
        "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    from ap_vast_unofficial_b200 import _capi as capi
    from ap_vast_unofficial_b200 import apvast
    from ap_vast_unofficial_b200.sharded import block_ranges
    from ap_vast_unofficial_b200.workloads import make_workload

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    else:
        torch.cuda.set_device(local_rank)
    lib = capi.lib()
    K, W = args.steps, args.warmup
    # each rank owns a contiguous block range of the 60-s signal; it only needs its own samples
    full = make_workload(args.workload, n_blocks=None)
    nblk_total = full["n_blocks"]
    sh = full["shapes"]
    H = sh["H"]
    t0, t1 = block_ranges(nblk_total, world)[rank]
    need = W + 2 * K + 3
    start = min(t0, max(0, nblk_total - need))
    sigA = full["signal_A"][start * H:(start + need) * H]
    sigB = full["signal_B"][start * H:(start + need) * H]
    np.random.seed(0)
    eng = apvast(rir_A=full["rir_A"], rir_B=full["rir_B"], perceptual=False, device=local_rank, **full["cfg"])
    V, L, n = sh["V"], sh["L"], sh["n"]

    # device-resident inputs for the kernel-only measurement
    d_sig = torch.from_numpy(np.stack([sigA, sigB])).cuda()
    torch.cuda.synchronize()

    def dev_ptr(sig, blk):
        return C.c_void_p(d_sig.data_ptr() + (sig * d_sig.shape[1] + blk * H) * 8)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    blk = 0
    # ---- warm-up (untimed): fills the statistics buffers and warms clocks / instruction caches
    for _ in range(W):
        capi.check(lib.apv_process_block_device(eng._h, dev_ptr(0, blk), dev_ptr(1, blk)))
        blk += 1
    capi.check(lib.apv_synchronize(eng._h))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: K steps, inputs resident in HBM, CUDA events on the engine's stream
    barrier()
    capi.check(lib.apv_timer_start(eng._h))
    launches = 0
    kt_panel, kt_syrk = 0.0, 0.0
    stage_acc = {}
    for _ in range(K):
        capi.check(lib.apv_process_block_device(eng._h, dev_ptr(0, blk), dev_ptr(1, blk)))
        blk += 1
        launches += int(lib.apv_launch_count(eng._h))
    ms = C.c_float(0)
    capi.check(lib.apv_timer_stop(eng._h, C.byref(ms)))
    barrier()
    dev_ms = float(ms.value)
    # per-kernel times of the LAST timed block (events recorded inside the timed region)
    kt = (C.c_float * 4)()
    capi.check(lib.apv_kernel_times(eng._h, kt))
    kt_panel, kt_syrk, n_panel = float(kt[0]), float(kt[1]), int(kt[2])
    stage_acc = eng.stage_times()

    # ---- timed region 2: end to end through the drop-in call, host buffers in, host buffers out
    # (one untimed call first: the host path's pinned staging buffer is allocated on first use)
    eng.process_input_buffers(sigA[blk * H:(blk + 1) * H], sigB[blk * H:(blk + 1) * H])
    blk += 1
    barrier()
    te0 = time.perf_counter()
    chk = 0.0
    trace = []
    for _ in range(K):
        tc0 = time.perf_counter()
        oA, oB, oAt, oBt = eng.process_input_buffers(sigA[blk * H:(blk + 1) * H], sigB[blk * H:(blk + 1) * H])
        chk += float(oA[0][0, 0])
        blk += 1
        trace.append(1e3 * (time.perf_counter() - tc0))
    if os.environ.get("APV_BENCH_TRACE"):
        print("e2e per-call ms:", " ".join("%.1f" % x for x in trace), file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- alternative configuration (not the headline): statistics by the structured evaluation (stats_mode=2)
    alt = None
    if not args.no_alt:
        np.random.seed(0)
        eng2 = apvast(rir_A=full["rir_A"], rir_B=full["rir_B"], perceptual=False, device=local_rank, stats_mode=2,
                      **full["cfg"])
        b2 = 0
        for _ in range(W):
            capi.check(lib.apv_process_block_device(eng2._h, dev_ptr(0, b2), dev_ptr(1, b2)))
            b2 += 1
        capi.check(lib.apv_synchronize(eng2._h))
        barrier()
        capi.check(lib.apv_timer_start(eng2._h))
        for _ in range(K):
            capi.check(lib.apv_process_block_device(eng2._h, dev_ptr(0, b2), dev_ptr(1, b2)))
            b2 += 1
        ms2 = C.c_float(0)
        capi.check(lib.apv_timer_stop(eng2._h, C.byref(ms2)))
        barrier()
        alt_ms = torch.tensor([float(ms2.value)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(alt_ms, op=dist.ReduceOp.MAX)
        st2 = eng2.stage_times()
        alt = {"what": "same workload with stats_mode=2 (first-row correlations + double-double diagonal recurrence "
                       "instead of the DMMA SYRK; identical parity, ~J/2 x fewer flops); device-timed like `value`",
               "value": world * K / (float(alt_ms[0].item()) * 1e-3), "unit": "updates/s",
               "ms_per_step": float(alt_ms[0].item()) / K, "S4_stats_ms": st2["S4_stats"]}
        eng2.close()

    t_dev = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t_dev[0].item()), float(t_dev[1].item())

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        nz = 2
        # algorithmic bytes of the tridiagonalisation per block: the trailing matrix is read once per column
        # (n^3/3 * 8 B per zone), both zones
        jj = np.arange(n - 1, dtype=np.float64)
        td_bytes = float(np.sum((n - jj - 1) ** 2) * 8.0 * nz)
        ach = td_bytes / (kt_panel * 1e-3) / 1e9 if kt_panel > 0 else 0.0
        tf = C.c_double(0)
        capi.check(lib.apv_bench_dmma_peak(4000, C.byref(tf)))
        M_, P_ = sh["M"], sh["N"] - sh["J"]
        syrk_flops = 4.0 * M_ * P_ * n * (n + 1)            # SYRK lower triangle, mul+add, 4 matrices
        ach_tf = syrk_flops / (kt_syrk * 1e-3) / 1e12 if kt_syrk > 0 else 0.0
        ups = world * K / (dev_ms_max * 1e-3)
        e2e = world * K / (e2e_ms_max * 1e-3)
        two_stage = n_panel < 0
        rl_stats = {"kernel": "syrk_toeplitz_kernel (FP64 DMMA statistics, implicit Toeplitz operand)", "bound": "tensor",
                    "achieved": ach_tf, "peak": float(tf.value), "unit": "TFLOP/s",
                    "frac": ach_tf / float(tf.value) if tf.value else None,
                    "peak_source": "FP64 mma.sync m8n8k4 issue-rate microbenchmark run in this process "
                                   "(MEASURED_PEAKS.json holds HBM and bf16 only)",
                    "algorithmic_flops_per_block": syrk_flops, "kernel_ms_per_block": kt_syrk,
                    "share_of_step": kt_syrk / (dev_ms_max / K) if dev_ms_max > 0 else None,
                    "traffic": 1.029e9 * 4,
                    "traffic_source": "profiles/r01_ncu_full_v3.txt (ncu --set full, one of the 4 launches per block: "
                                      "7.7 MB read + 1.02 GB written -- the per-microphone partial matrices)",
                    "ncu_tensor_pipe_pct": 94.4}
        if two_stage:
            td_flops = 2.0 * 4.0 * n ** 3 / 3.0           # 4 n^3 / 3 per zone
            td_ms = float(stage_acc.get("S5_tridiag", 0.0))
            rl_td = {"kernel": "two-stage tridiagonalisation (band.cu: sb_panel_qr + DMMA gemm | sb2st_chase)",
                     "bound": "tensor", "achieved": td_flops / (td_ms * 1e-3) / 1e12 if td_ms > 0 else 0.0,
                     "peak": float(tf.value), "unit": "TFLOP/s",
                     "frac": td_flops / (td_ms * 1e-3) / 1e12 / float(tf.value) if td_ms > 0 and tf.value else None,
                     "algorithmic_flops_per_block": td_flops, "ms_per_block": td_ms,
                     "dense_to_band_ms": kt_panel, "band_to_tridiagonal_ms": float(kt[3]),
                     "note": "latency-bound stages (cluster QR columns, bulge-chasing steps) beside the DMMA GEMMs"}
            roof, roof_other = rl_stats, {"roofline_tridiag": rl_td}
        else:
            rl_td = {"kernel": "td_panel_kernel (Householder tridiagonalisation, both zones)", "bound": "hbm",
                     "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_block": td_bytes,
                     "launches_per_block": n_panel, "kernel_ms_per_block": kt_panel, "traffic": None,
                     "traffic_sample": {"source": "profiles/r01_ncu_full_v2.txt (ncu --set full, panel 5 of 128, cfg3)",
                                        "dram_bytes": 8.165e9, "algorithmic_bytes": 7.874e9, "ratio": 1.04}}
            roof, roof_other = (rl_td, {"roofline_stats": rl_stats}) if kt_panel > kt_syrk else (rl_stats, {"roofline_tridiag": rl_td})
        line = {
            "metric": "filter_updates_per_sec", "value": ups, "unit": "updates/s", "rtf": ups * H / FS,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: synthetic 2-zone L={sh['L']} M={sh['M']} J={sh['J']} n={n} "
                                   f"K={sh['K']} Nb={sh['Nb']} H={H} N={sh['N']} V={V} fs=48000",
                       "sharding": f"contiguous block ranges over {world} rank(s), no per-block collective",
                       "l2": "per-block working set ~1.6 GB (4 R + jdiag workspace) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": "updates/s", "rtf": e2e * H / FS, "ms_per_step": e2e_ms_max / K,
                    "h2d_bytes_per_step": 2 * H * 8, "d2h_bytes_per_step": (2 * V * H * L + 2 * H) * 8 + 32},
            "gpu_launches": launches,
            "roofline": roof,
            "stage_ms_last_block": stage_acc, "clocks": clocks, "checksum": chk,
        }
        line.update(roof_other)
        line["alt_structured_stats"] = alt
        if world == 1 and not args.no_cpu_baseline:
            wl = make_workload(args.workload, n_blocks=8)
            times, split, sample = cpu_reference_sample(wl, 1, 3)
            line["cpu_baseline"] = {"value": 1.0 / float(np.mean(times)), "unit": "updates/s", "cores": blas_threads(),
                                    "kind": "port", "sample": sample, "split_s": split, "host_cpus": os.cpu_count()}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--workload", default="cfg3", choices=["cfg2", "cfg3", "small"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra measurement of the structured-statistics mode")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
