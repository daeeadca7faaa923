"""torchrun --nproc-per-node N scripts/run_sharded_nccl.py [workload] [nblocks]
Block-range sharding with the CUDA engine over NCCL: every rank processes its contiguous block range, the
overlap-add tail goes to the right neighbour (isend/irecv), filters are all-gathered; rank 0 also runs the
single-stream job and reports the worst deviation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ap_vast_unofficial_b200 import apvast
from ap_vast_unofficial_b200.sharded import process_signal_sharded
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
    res = process_signal_sharded(make, wl["signal_A"], wl["signal_B"], rank=rank, world=world, dist=dist, seed=0)
    dist.barrier(); dt = time.perf_counter() - t0
    # gather every rank's outputs on rank 0 for the check
    outs = [None] * world
    dist.all_gather_object(outs, (res["blocks"], res["out_A"]))
    if rank == 0:
        ref = process_signal_sharded(make, wl["signal_A"], wl["signal_B"], seed=0)
        worst_o = worst_w = 0.0
        for blocks, oa in outs:
            for k, t in enumerate(range(*blocks)):
                worst_o = max(worst_o, float(np.linalg.norm(oa[k] - ref["out_A"][t]) / np.linalg.norm(ref["out_A"][t])))
        for t, w in enumerate(res["all_w_A"]):
            for v in range(w.shape[0]):
                worst_w = max(worst_w, float(np.linalg.norm(w[v] - ref["w_A"][t][v]) / np.linalg.norm(ref["w_A"][t][v])))
        print(f"sharded NCCL run: world={world} workload={name} blocks={nb} wall={dt:.2f}s  "
              f"worst output deviation {worst_o:.2e}  worst filter deviation {worst_w:.2e} vs single stream")
        assert worst_o < 1e-8 and worst_w < 1e-8
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
