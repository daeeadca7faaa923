"""Block-range sharding of one long signal over the GPUs of a box.

The filter at block t is a finite-memory function of the inputs, so rank g of G can own the contiguous
block range [t_g, t_{g+1}) and reproduce the single-stream result exactly:

* it replays the cheap state stages S1-S3 (``advance_state``) over a halo of
  ``warmup_blocks = ceil(N/H) + 1 + ceil((K-1)/H)`` blocks before t_g (or starts at block 0 with the
  reference's seeded start buffers when the halo reaches the beginning);
* the only data that crosses a boundary is the output overlap-add tail ``G[:, H:, :]`` written by the
  last block of rank g-1 with *that* block's filter: it is sent to rank g (one point-to-point message,
  NCCL over NVLink on GPUs, gloo in the CPU tests) and added to the first output blocks of rank g;
* the per-rank outputs and filters are gathered on rank 0 at the end.

No collective is on the per-block path.  ``comm`` is ``torch.distributed`` (already initialised) or None.
"""
from __future__ import annotations

import math

import numpy as np


def block_ranges(n_blocks: int, world: int):
    """Contiguous, balanced block ranges [(t0, t1), ...] for `world` ranks."""
    base, rem = divmod(n_blocks, world)
    out, t = [], 0
    for g in range(world):
        c = base + (1 if g < rem else 0)
        out.append((t, t + c))
        t += c
    return out


def warmup_blocks(stats_len: int, hop: int, rir_len: int) -> int:
    return math.ceil(stats_len / hop) + 1 + math.ceil((rir_len - 1) / hop)


def _to_dev(t, dist):
    import torch
    if dist.get_backend() == "nccl":
        return t.cuda()
    return t


def process_signal_sharded(make_engine, signal_A, signal_B, rank: int = 0, world: int = 1, dist=None, seed=0,
                           keep_outputs=True):
    """Run the whole signal, block-range sharded.  Returns on every rank a dict with this rank's
    ``blocks`` (t0, t1), ``w_A``/``w_B`` lists (one (V, n) array per owned block) and ``out_A``/``out_B``
    lists ((V, H, L) per owned block, halo already applied); on rank 0 additionally ``all_w_A``/``all_w_B``
    (gathered filters for every block, in block order).

    ``make_engine()`` must return an object with the reference interface (``process_input_buffers``,
    ``hop_size``, ``w_A``, ``w_B``, ``output_A_overlap_buffer`` ...) and optionally ``advance_state``.
    """
    np.random.seed(seed)                       # rank 0 reproduces the reference's randn start exactly
    eng = make_engine()
    H = eng.hop_size
    n_blocks = len(signal_A) // H
    t0, t1 = block_ranges(n_blocks, world)[rank]
    wu = warmup_blocks(eng.statistics_buffer_length, H, eng.rir_length)
    start = max(0, t0 - wu)
    adv = getattr(eng, "advance_state", None)
    for t in range(start, t0):
        a, b = signal_A[t * H:(t + 1) * H], signal_B[t * H:(t + 1) * H]
        if adv is not None:
            adv(a, b)
        else:                                   # engines without a state-only call run the full block
            eng.process_input_buffers(a, b)
    if adv is None and t0 > start:
        # a full-block warm-up leaves its own overlap tail in the output buffers; it is part of the exact
        # single-stream result, so no halo is needed from the left neighbour for this engine
        need_halo = False
    else:
        need_halo = t0 > 0
    res = dict(blocks=(t0, t1), w_A=[], w_B=[], out_A=[], out_B=[])
    for t in range(t0, t1):
        oA, oB, _, _ = eng.process_input_buffers(signal_A[t * H:(t + 1) * H], signal_B[t * H:(t + 1) * H])
        wA, wB = eng.w_A, eng.w_B
        res["w_A"].append(None if wA is None else np.array(wA[:, :, 0]))
        res["w_B"].append(None if wB is None else np.array(wB[:, :, 0]))
        if keep_outputs:
            res["out_A"].append(None if oA is None else np.stack(oA))
            res["out_B"].append(None if oB is None else np.stack(oB))
    # ---- overlap-add halo: tail of the last owned block -> right neighbour
    Nb = eng.block_size
    if world > 1 and dist is not None:
        import torch
        tails = []
        for nm in ("output_A_overlap_buffer", "output_B_overlap_buffer"):
            g = getattr(eng, nm, None)
            tails.append(np.zeros((eng.number_of_eigenvectors, Nb - H, eng.number_of_srcs)) if g is None or t1 == t0
                         else np.array(g[:, H:, :]))
        send = torch.from_numpy(np.ascontiguousarray(np.stack(tails)))
        recv = torch.zeros_like(send)
        reqs = []
        if rank + 1 < world:
            reqs.append(dist.isend(_to_dev(send, dist), rank + 1))
        rbuf = None
        if rank > 0:
            rbuf = _to_dev(recv, dist)
            reqs.append(dist.irecv(rbuf, rank - 1))
        for r in reqs:
            r.wait()
        if rank > 0 and need_halo and keep_outputs and t1 > t0:
            tail = rbuf.cpu().numpy()
            for z, key in enumerate(("out_A", "out_B")):
                for k in range(math.ceil((Nb - H) / H)):
                    if k < len(res[key]) and res[key][k] is not None:
                        seg = tail[z][:, k * H:(k + 1) * H, :]
                        res[key][k][:, :seg.shape[1], :] += seg
        # ---- final gather of the filters on rank 0
        for key in ("w_A", "w_B"):
            mine = [w for w in res[key] if w is not None]
            gathered = [None] * world
            dist.all_gather_object(gathered, mine) if dist.get_backend() != "nccl" else _gather_nccl(dist, mine, gathered, eng)
            if rank == 0:
                res["all_" + key] = [w for part in gathered for w in part]
    else:
        res["all_w_A"], res["all_w_B"] = res["w_A"], res["w_B"]
    return res


def _gather_nccl(dist, mine, gathered, eng):
    """all_gather of variable-length filter lists over NCCL (pad to the longest range)."""
    import torch
    world = dist.get_world_size()
    V, n = eng.number_of_eigenvectors, eng.filter_length * eng.number_of_srcs
    cnt = torch.tensor([len(mine)], device="cuda")
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    mx = max(int(c.item()) for c in cnts)
    buf = torch.zeros((max(mx, 1), V, n), dtype=torch.float64, device="cuda")
    if mine:
        buf[:len(mine)] = torch.from_numpy(np.stack(mine)).cuda()
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    for g in range(world):
        k = int(cnts[g].item())
        gathered[g] = [bufs[g][i].cpu().numpy() for i in range(k)]
