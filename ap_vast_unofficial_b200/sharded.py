"""Block-range sharding of one long signal over the GPUs of a box.

The filter at block t is a finite-memory function of the inputs (the streaming state of reference
``Python/apvast.py:115-151``), so rank g of G can own the contiguous block range [t_g, t_{g+1}) and reproduce the
single-stream result exactly:

* it replays the cheap state stages S1-S3 over a halo of ``warmup_blocks`` blocks before t_g (or starts at block 0
  with the reference's seeded start buffers when the halo reaches the beginning);
* the only data that crosses a boundary is the output overlap-add tail ``G[:, H:, :]`` (``apvast.py:455-465``) written
  by the last ``Nb/H - 1`` blocks of rank g-1: it is sent to rank g (one point-to-point message) and added to the
  first ``Nb/H - 1`` output blocks of rank g -- every range but the last must therefore hold at least that many blocks;
* outputs and filters are gathered on rank 0 at the end (nobody else needs them).

Two drivers:

``RangeRunner``            the product path.  Everything stays on the device behind the C-ABI (``apv_range_run``,
                           ``apv_range_exchange_halo``, ``apv_range_gather``): the tail goes HBM -> NVLink -> HBM with
                           ``ncclSend``/``ncclRecv`` issued by the library itself on the engine's stream, the gather
                           lands in rank 0's HBM and leaves through one D2H copy.  ``torch.distributed`` is used once,
                           to hand the NCCL unique id to the other ranks.
``process_signal_sharded`` the same protocol written against the reference's Python interface with
                           ``torch.distributed`` point-to-point calls; it runs any engine that has the reference's
                           constructor / per-block call (the oracle on CPU over gloo in ``tests/test_sharded_gloo.py``).
"""
from __future__ import annotations

import ctypes as C
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


def warmup_blocks(stats_len: int, hop: int, rir_len: int, block_size: int = None) -> int:
    """Blocks of S1-S3 a late-started engine needs before its statistics equal the single stream's.

    The statistics buffer holds ceil(N/H) appended hops; each appended hop sums Nb/H overlapping WOLA frames; each
    frame reads a response buffer of Nb/H hops of FIR output; each FIR output reaches ceil((K-1)/H) hops back:
    ceil(N/H) - 1 + 2 (Nb/H - 1) + ceil((K-1)/H).  For hop = Nb/2 that is the ceil(N/H) + 1 + ceil((K-1)/H) SURVEY.md
    section 5 measured (4 blocks at cfg-1)."""
    k = 2 if block_size is None else math.ceil(block_size / hop)
    return math.ceil(stats_len / hop) - 1 + 2 * (k - 1) + math.ceil((rir_len - 1) / hop)


def tail_blocks(block_size: int, hop: int) -> int:
    """Output blocks the overlap-add tail of a range reaches into (= minimum size of every range but the last)."""
    return math.ceil(block_size / hop) - 1


# ------------------------------------------------------------------------------------------------ device path
class RangeRunner:
    """One rank of a block-range-sharded run, on the device (C-ABI ``apv_range_*``).

    ``eng`` is an ``ap_vast_unofficial_b200.apvast``; ``dist`` an initialised ``torch.distributed`` (any backend) or
    None for a single rank.  ``max_owned`` / ``max_halo`` size the device buffers once (``apv_range_reserve``)."""

    def __init__(self, eng, rank: int = 0, world: int = 1, dist=None, max_owned: int = 1, max_halo: int = None,
                 total_blocks: int = None):
        from . import _capi as capi
        self.capi, self.lib, self.eng = capi, capi.lib(), eng
        self.rank, self.world = rank, world
        self.halo = warmup_blocks(eng.statistics_buffer_length, eng.hop_size, eng.rir_length, eng.block_size)
        self.k1 = tail_blocks(eng.block_size, eng.hop_size)
        V, H, L, n = eng.number_of_eigenvectors, eng.hop_size, eng.number_of_srcs, eng.filter_length * eng.number_of_srcs
        self.per_out, self.per_w = 2 * V * H * L, 2 * V * n
        if world > 1:
            if dist is None:
                raise RuntimeError("RangeRunner: world > 1 needs torch.distributed to hand out the NCCL unique id")
            uid = (C.c_ubyte * 128)()
            if rank == 0:
                capi.check(self.lib.apv_comm_unique_id(uid))
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=0)
            uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
            capi.check(self.lib.apv_comm_init(eng._h, rank, world, uid))
        else:
            capi.check(self.lib.apv_comm_init(eng._h, 0, 1, None))
        total = (max_owned * world if total_blocks is None else total_blocks) if rank == 0 else 0
        capi.check(self.lib.apv_range_reserve(eng._h, self.halo if max_halo is None else max_halo, max(max_owned, 1), total))
        self.bytes_halo = 2 * V * L * (eng.block_size - H) * 8           # one tail message
        self.owned = 0

    def run(self, sig_A, sig_B, n_halo: int, n_owned: int, device_ptrs=None):
        """Enqueue n_halo state-only blocks + n_owned full blocks.  ``sig_*``: host arrays of (n_halo + n_owned) * H
        samples; or pass ``device_ptrs=(ptr_A, ptr_B)`` for hops already resident in HBM.  Asynchronous."""
        if self.rank + 1 < self.world and n_owned < self.k1:
            raise RuntimeError(f"a block range must hold at least Nb/H - 1 = {self.k1} blocks")
        if device_ptrs is not None:
            pa, pb, dev = C.c_void_p(device_ptrs[0]), C.c_void_p(device_ptrs[1]), 1
        else:
            a = np.ascontiguousarray(sig_A, dtype=np.float64).reshape(-1)
            b = np.ascontiguousarray(sig_B, dtype=np.float64).reshape(-1)
            if a.size != (n_halo + n_owned) * self.eng.hop_size or b.size != a.size:
                raise RuntimeError("invalid input size")
            self._keep = (a, b)
            pa, pb, dev = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 0
        self.eng._sync_flags()
        self.capi.check(self.lib.apv_range_run(self.eng._h, n_halo, n_owned, pa, pb, dev))
        self.eng._blocks += n_owned
        self.owned = n_owned

    def exchange_halo(self):
        """Overlap-add tail of the left neighbour -> my first Nb/H - 1 output blocks (device to device, NCCL)."""
        self.capi.check(self.lib.apv_range_exchange_halo(self.eng._h))

    def gather(self, counts, out_host=None, w_host=None, root: int = 0):
        """Outputs (total, 2, V, H, L) and filters (total, 2, V, n) of all ranks on `root` (host arrays, ideally
        pinned: ``_capi.pinned_array``); other ranks pass nothing.  Synchronises."""
        cnt = (C.c_int * self.world)(*[int(c) for c in counts])
        po = None if out_host is None else out_host.ctypes.data_as(C.c_void_p)
        pw = None if w_host is None else w_host.ctypes.data_as(C.c_void_p)
        self.capi.check(self.lib.apv_range_gather(self.eng._h, root, cnt, po, pw))

    def close(self):
        self.capi.check(self.lib.apv_comm_destroy(self.eng._h))


def process_signal_device(make_engine, signal_A, signal_B, rank: int = 0, world: int = 1, dist=None, seed=0):
    """Whole signal, block-range sharded, device path.  Returns on rank 0 ``(out, w)``: outputs
    (n_blocks, 2, V, H, L) and filters (n_blocks, 2, V, n) of the whole signal in block order; ``(None, None)``
    elsewhere."""
    np.random.seed(seed)                       # rank 0 reproduces the reference's randn start exactly
    eng = make_engine()
    H = eng.hop_size
    n_blocks = len(signal_A) // H
    ranges = block_ranges(n_blocks, world)
    t0, t1 = ranges[rank]
    rr = RangeRunner(eng, rank, world, dist, max_owned=max(t1 - t0, 1), total_blocks=n_blocks)
    start = max(0, t0 - rr.halo)
    rr.run(signal_A[start * H:t1 * H], signal_B[start * H:t1 * H], t0 - start, t1 - t0)
    rr.exchange_halo()
    out = w = None
    if rank == 0:
        out = np.empty((n_blocks, 2, eng.number_of_eigenvectors, H, eng.number_of_srcs))
        w = np.empty((n_blocks, 2, eng.number_of_eigenvectors, eng.filter_length * eng.number_of_srcs))
    rr.gather([b - a for a, b in ranges], out, w)
    rr.close()
    return out, w


# ------------------------------------------------------------------------------------------------ generic path
def process_signal_sharded(make_engine, signal_A, signal_B, rank: int = 0, world: int = 1, dist=None, seed=0,
                           keep_outputs=True, gather=True):
    """Run the whole signal, block-range sharded, through the reference's Python interface.  Returns on every rank
    a dict with this rank's ``blocks`` (t0, t1), ``w_A``/``w_B`` lists (one (V, n) array per owned block) and
    ``out_A``/``out_B`` lists ((V, H, L) per owned block, halo already applied); on rank 0 additionally
    ``all_w_A``/``all_w_B`` (gathered filters for every block, in block order).

    ``make_engine()`` must return an object with the reference interface (``process_input_buffers``,
    ``hop_size``, ``w_A``, ``w_B``, ``output_A_overlap_buffer`` ...) and optionally ``advance_state``.
    """
    np.random.seed(seed)                       # rank 0 reproduces the reference's randn start exactly
    eng = make_engine()
    H = eng.hop_size
    Nb = eng.block_size
    n_blocks = len(signal_A) // H
    t0, t1 = block_ranges(n_blocks, world)[rank]
    wu = warmup_blocks(eng.statistics_buffer_length, H, eng.rir_length, Nb)
    k1 = tail_blocks(Nb, H)
    if world > 1 and rank + 1 < world and t1 - t0 < k1:
        raise RuntimeError(f"a block range must hold at least Nb/H - 1 = {k1} blocks")
    start = max(0, t0 - wu)
    adv = getattr(eng, "advance_state", None)
    for t in range(start, t0):
        a, b = signal_A[t * H:(t + 1) * H], signal_B[t * H:(t + 1) * H]
        if adv is not None:
            adv(a, b)
        else:                                   # engines without a state-only call run the full block
            eng.process_input_buffers(a, b)
    # a full-block warm-up (engines without a state-only call) leaves its own overlap tail in the output buffers; the
    # warm-up is longer than Nb/H - 1 blocks, so that tail is the exact single-stream one and no halo is needed
    need_halo = adv is not None and t0 > 0
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
    if world > 1 and dist is not None:
        import torch
        nccl = dist.get_backend() == "nccl"
        V, L = eng.number_of_eigenvectors, eng.number_of_srcs
        tails = []
        for nm in ("output_A_overlap_buffer", "output_B_overlap_buffer"):
            g = getattr(eng, nm, None)
            tails.append(np.zeros((V, Nb - H, L)) if g is None or t1 == t0 else np.array(g[:, H:, :]))
        send = torch.from_numpy(np.ascontiguousarray(np.stack(tails)))
        recv = torch.zeros_like(send)
        if nccl:
            send, recv = send.cuda(), recv.cuda()
        reqs = []
        if rank + 1 < world:
            reqs.append(dist.isend(send, rank + 1))
        if rank > 0:
            reqs.append(dist.irecv(recv, rank - 1))
        for r in reqs:
            r.wait()
        if rank > 0 and need_halo and keep_outputs and t1 > t0:
            tail = recv.cpu().numpy()
            for z, key in enumerate(("out_A", "out_B")):
                for k in range(k1):
                    if k < len(res[key]) and res[key][k] is not None:
                        seg = tail[z][:, k * H:(k + 1) * H, :]
                        res[key][k][:, :seg.shape[1], :] += seg
        # ---- final gather of the filters: to rank 0 only (point to point; nobody else needs them)
        for key in ("w_A", "w_B") if gather else ():
            mine = [w for w in res[key] if w is not None]
            if rank == 0:
                parts = [mine]
                for src in range(1, world):
                    cnt = torch.zeros(1, dtype=torch.int64)
                    cnt = cnt.cuda() if nccl else cnt
                    dist.recv(cnt, src)
                    k = int(cnt.item())
                    buf = torch.zeros((k, V, eng.filter_length * L), dtype=torch.float64)
                    buf = buf.cuda() if nccl else buf
                    if k:
                        dist.recv(buf, src)
                    parts.append([buf[i].cpu().numpy() for i in range(k)])
                res["all_" + key] = [w for part in parts for w in part]
            else:
                cnt = torch.tensor([len(mine)], dtype=torch.int64)
                dist.send(cnt.cuda() if nccl else cnt, 0)
                if mine:
                    buf = torch.from_numpy(np.stack(mine))
                    dist.send(buf.cuda() if nccl else buf, 0)
    else:
        res["all_w_A"], res["all_w_B"] = res["w_A"], res["w_B"]
    return res
