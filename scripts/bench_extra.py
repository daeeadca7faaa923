"""bench.py --workload cfg4 | cfg5: the two BASELINE configs that are not block streams of the two-zone engine.

cfg4  mu x V trade-off sweep at L=16, J=256 (n = 4096): 8 mu values x the FULL rank range (V = n) from ONE joint
      diagonalisation per zone and block (the reference redoes jdiag per mu, apvast.py:378-382).  A step = one block
      with V = n (S1..S7, all 4096 ranks rendered) + the sweep: filters for the 8 mu into an HBM buffer
      (8 x 2 x 4096 x 4096 doubles = 2.1 GB, they stay on the device) + the eigen-basis figures of merit to the host.
      The full-rank filter is checked against the closed form w = (R_B + mu (R_D + reg I))^-1 r_B once, untimed.
cfg5  L=32, J=256 (n = 8192), 4 zones with per-zone perceptual weighting, 64 independent clips on 8 GPUs: rank g runs
      its 8 clips one after the other (one 4-zone engine set, state reset per clip), no collective on the data path
      (clips are independent: "replicas only" plus the final reduction of the timings).  A step = one 4-zone update.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np

FS = 48000.0


def run_cfg4(args, rank, world, local_rank):
    import torch
    from ap_vast_unofficial_b200 import _capi as capi, apvast
    from ap_vast_unofficial_b200.workloads import make_workload
    torch.cuda.set_device(local_rank)
    K, W = args.steps, args.warmup
    wl = make_workload("cfg3", n_blocks=W + K + 1)
    sh = wl["shapes"]
    n, H, L = sh["n"], sh["H"], sh["L"]
    cfg = dict(wl["cfg"]); cfg["number_of_eigenvectors"] = n
    mus = np.logspace(-3, 1, 8)
    np.random.seed(0)
    eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, device=local_rank, **cfg)
    lib = capi.lib()
    d_sig = torch.from_numpy(np.stack([wl["signal_A"], wl["signal_B"]])).cuda()
    d_w = torch.empty((len(mus), 2, n, n), dtype=torch.float64, device="cuda")
    pA, pB = d_sig.data_ptr(), d_sig.data_ptr() + d_sig.shape[1] * 8

    def step(b):
        capi.check(lib.apv_process_block_device(eng._h, C.c_void_p(pA + b * H * 8), C.c_void_p(pB + b * H * 8)))
        return eng.sweep_metrics(mus, device_out=d_w.data_ptr())

    for b in range(W):
        step(b)
    torch.cuda.synchronize()
    capi.check(lib.apv_timer_start(eng._h))
    t0 = time.perf_counter()
    for b in range(W, W + K):
        met = step(b)
    ms = C.c_float(0)
    capi.check(lib.apv_timer_stop(eng._h, C.byref(ms)))
    wall = time.perf_counter() - t0
    st = eng.stage_times()
    # closed form at full rank (untimed): w[V-1] = (R_B + mu (R_D + reg I))^-1 r_B
    errs = []
    for zi, (RB, RD, r) in enumerate(((eng.R_A_to_A, eng.R_A_to_B, eng.r_A[:, 0]), (eng.R_B_to_B, eng.R_B_to_A, eng.r_B[:, 0]))):
        for k in (0, 4, 7):
            closed = np.linalg.solve(RB + mus[k] * (RD + 1e-7 * np.eye(n)), r)
            got = d_w[k, zi, n - 1].cpu().numpy()
            errs.append(float(np.linalg.norm(got - closed) / np.linalg.norm(closed)))
    ups = K / (float(ms.value) * 1e-3)
    line = {"metric": "filter_updates_per_sec", "value": ups, "unit": "updates/s", "rtf": ups * H / FS, "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": float(ms.value) / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg4: mu x V sweep, 8 mu x full rank range V = n = {n} (L={L} J={sh['J']}), one joint "
                                   f"diagonalisation per zone and block; filters (8, 2, {n}, {n}) stay in HBM"},
            "e2e": {"value": K / wall, "unit": "updates/s", "ms_per_step": 1e3 * wall / K,
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(met.size * 8),
                    "note": "hops resident in HBM; per step the eigen-basis figures of merit (8, 2, V, 3) go to the host"},
            "gpu_launches": int(lib.apv_launch_count(eng._h)) * K,
            "stage_ms_last_block": st, "closed_form_rel_l2": errs, "closed_form_bar": 1e-10,
            "figures_of_merit_last_block": {"mu": mus.tolist(), "dark_energy_full_rank_A": met[:, 0, -1, 0].tolist(),
                                            "bright_energy_full_rank_A": met[:, 0, -1, 1].tolist()}}
    print(json.dumps(line), flush=True)
    assert max(errs) < 1e-10, errs
    eng.close()


def make_cfg5(seed=0, n_zones=4, L=32, M=16, J=256, K=1024):
    from ap_vast_unofficial_b200.workloads import _rirs
    rirs = [_rirs(50 + 7 * z + 1000 * seed, K, L, M) for z in range(n_zones)]
    return rirs


def run_cfg5(args, rank, world, local_rank):
    import torch
    from ap_vast_unofficial_b200.workloads import _programme
    from ap_vast_unofficial_b200.zones import apvast_zones
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    Z, L, M, J, Kr, Nb, N, V = 4, 32, 16, 256, 1024, 2048, 2048, 64
    H = Nb // 2
    clips_total = int(os.environ.get("APV_CFG5_CLIPS", "64"))
    per_rank = max(1, clips_total // world)
    blocks = args.steps                   # timed 4-zone updates per clip
    rirs = make_cfg5()
    np.random.seed(0)
    eng = apvast_zones(block_size=Nb, rirs=rirs, filter_length=J, modeling_delay=32, reference_indices=[0, 1, 2, 3],
                       number_of_eigenvectors=V, mu=1.0, statistics_buffer_length=N, perceptual=not args.no_perceptual,
                       device=local_rank)
    states = [e.get_state() for e in eng.engines]
    sig0 = [_programme(900 + z, (blocks + 1) * H) for z in range(Z)]
    eng.process_input_buffers([s[:H] for s in sig0])          # warm-up of kernels and clocks, untimed
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    done = 0
    chk = 0.0
    for c in range(per_rank):
        clip = rank * per_rank + c
        for e, s in zip(eng.engines, states):                 # every clip starts from the same fresh state
            e.set_state(s)
        sig = [_programme(1000 + 10 * clip + z, blocks * H) for z in range(Z)]
        for b in range(blocks):
            outs = eng.process_input_buffers([s[b * H:(b + 1) * H] for s in sig])
            done += 1
        chk += float(outs[0][0][0, 0])
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    if rank == 0:
        ups = world * done / sec
        st = eng.stage_times()
        line = {"metric": "filter_updates_per_sec", "value": ups, "unit": "4-zone updates/s", "rtf": ups * H / FS,
                "n_gpus": world, "steps": blocks, "warmup": 1, "ms_per_step": 1e3 * sec / done, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"cfg5: {Z} zones, L={L} J={J} n={L * J} M={M} per zone, per-zone perceptual weighting "
                                       f"{'on' if not args.no_perceptual else 'off'}, {per_rank * world} clips x {blocks} updates, "
                                       f"{per_rank} clips per GPU, no data-path collective (independent clips)"},
                "e2e": {"value": ups, "unit": "4-zone updates/s", "h2d_bytes_per_step": Z * 2 * H * 8,
                        "d2h_bytes_per_step": Z * (V * H * L + 2 * H * L) * 8,
                        "note": "host buffers in and out through zones.apvast_zones.process_input_buffers; wall clock, max over ranks"},
                "gpu_launches": sum(s["launches"] for s in st) * done,
                "stage_ms_last_update_per_zone": st, "checksum": chk}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run(args, rank, world, local_rank):
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable":
                              f"{args.workload}: the reference has no such mode (it redoes jdiag per mu / has exactly two zones)"}))
        return
    if args.workload == "cfg4":
        if rank == 0:
            run_cfg4(args, rank, 1, local_rank)
    else:
        run_cfg5(args, rank, world, local_rank)
