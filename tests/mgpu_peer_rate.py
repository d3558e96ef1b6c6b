"""Run under torchrun with 2+ GPUs (measurement aid, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29540 tests/mgpu_peer_rate.py [ipc|symm]

What does the sharded SGD step sustain on UNIFORM ids when the item rows are (a) all local, (b) uniform over all
ranks, (c) all remote -- with the tables mapped the way the product maps them (one process per GPU, CUDA IPC or
symmetric memory)?  profiles/peer_probe.cu answers the same question for a single process with
cudaDeviceEnablePeerAccess; the difference between the two is the mapping."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else "ipc"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = dist.get_rank(), dist.get_world_size()
    from neurec_b200 import ops
    from neurec_b200.util import peer
    dim, ni_l, nu_l, n = 128, 12_500_000, 1_000_000, 1 << 20
    VS = peer.alloc_sharded(ni_l, dim, backend)
    US = peer.alloc_sharded(nu_l, dim, backend)
    VS.local.normal_(0, 0.01); US.local.normal_(0, 0.01)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    users = (torch.randint(0, nu_l, (n,), device="cuda", generator=g) + rank * nu_l).to(torch.int32)

    def ids(kind):
        if kind == "local":
            lo, hi = rank * ni_l, (rank + 1) * ni_l
        elif kind == "remote":
            r = (rank + 1) % ws
            lo, hi = r * ni_l, (r + 1) * ni_l
        else:
            lo, hi = 0, ni_l * ws
        return tuple(torch.randint(lo, hi, (n,), device="cuda", generator=g).to(torch.int32) for _ in range(2))
    loss = torch.zeros(1, device="cuda")
    for kind in ("local", "uniform", "remote", "local"):
        pos, neg = ids(kind)
        for both in (True, False):                  # False: only rank 0 runs the kernel (its peers' SMs are idle)
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if both or rank == 0:
                ops.mf_bpr_sgd_sharded(US, VS, rank, users, pos, neg, 0.01, 0.0, loss)
                torch.cuda.synchronize()
                a.record()
                for _ in range(3):
                    ops.mf_bpr_sgd_sharded(US, VS, rank, users, pos, neg, 0.01, 0.0, loss)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
            else:
                ms = 0.0
            dist.barrier()
            if rank == 0:
                print("%s | item rows %-7s | %-22s | %8.3f ms per 2^20 triplets = %7.1f M triplets/s" % (
                    backend, kind, "all ranks at once" if both else "rank 0 alone", ms, n / ms / 1e3), flush=True)
    US.close(); VS.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
