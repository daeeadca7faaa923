"""torchrun --nproc-per-node N scripts/run_sharded_nccl.py [workload] [nblocks]
Block-range sharding with the CUDA engine on N GPUs, device path (sharded.process_signal_device: apv_range_run,
overlap-add tail over ncclSend/ncclRecv inside the library, gather into rank 0's HBM).  Rank 0 also runs the whole
signal as ONE stream through the per-hop call and reports the worst deviation of every block's outputs and filters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
from ap_vast_unofficial_b200 import _capi as capi, apvast
from ap_vast_unofficial_b200.sharded import process_signal_device
from ap_vast_unofficial_b200.workloads import make_workload


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    wl = make_workload(name, n_blocks=nb)
    make = lambda: apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, device=local, **wl["cfg"])
    dist.barrier(); t0 = time.perf_counter()
    out, w = process_signal_device(make, wl["signal_A"], wl["signal_B"], rank=rank, world=world, dist=dist, seed=0)
    dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        v = C.c_int(0)
        capi.check(capi.lib().apv_nccl_version(C.byref(v)))
        np.random.seed(0)
        eng = make()
        H = eng.hop_size
        worst_o = worst_w = 0.0
        for t in range(nb):
            oA, oB, _, _ = eng.process_input_buffers(wl["signal_A"][t * H:(t + 1) * H], wl["signal_B"][t * H:(t + 1) * H])
            ref_o = np.stack([np.stack(oA), np.stack(oB)])
            ref_w = np.stack([eng.w_A[:, :, 0], eng.w_B[:, :, 0]])
            worst_o = max(worst_o, float(np.linalg.norm(out[t] - ref_o) / np.linalg.norm(ref_o)))
            worst_w = max(worst_w, float(np.linalg.norm(w[t] - ref_w) / np.linalg.norm(ref_w)))
        print(f"sharded run, device path: world={world} workload={name} blocks={nb} wall={dt:.2f}s NCCL {v.value}  "
              f"worst output deviation {worst_o:.2e}  worst filter deviation {worst_w:.2e} vs single stream", flush=True)
        assert worst_o < 1e-8 and worst_w < 1e-8
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
