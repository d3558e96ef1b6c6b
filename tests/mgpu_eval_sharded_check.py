"""Run under torchrun with N >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29522 tests/mgpu_eval_sharded_check.py

Item-sharded evaluator: user AND item tables row-sharded over the ranks; per
batch ONE all-gather of the batch's user rows, the local candidate pass on every item shard, ONE
all-gather of the [B, K+1] (id, score) lists, merge + metrics.  Every rank checks that its merged
rows are bit-identical to the unsharded evaluator run on the full tables (which every rank can
build here because the check is small), then the throughput is reported."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = dist.get_rank(), dist.get_world_size()
    from neurec_b200 import ops
    from neurec_b200.evaluator import sharded
    from conftest import random_csr
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nu, ni, dim, K, B = 8192, 262144, 64, 20, 2048
    metric = ["Precision", "Recall", "NDCG", "MAP", "MRR"]
    rs = np.random.RandomState(1)                        # same on every rank
    U = (rs.randn(nu, dim) * 0.1).astype(np.float32); V = (rs.randn(ni, dim) * 0.1).astype(np.float32)
    tp, ti = random_csr(rs, nu, ni, np.full(nu, 50))
    sp, si = random_csr(rs, nu, ni, np.full(nu, 10))
    per_i = (ni + ws - 1) // ws
    lo, hi = rank * per_i, min(ni, (rank + 1) * per_i)
    lp, li = sharded.ItemShard.restrict_csr(tp, ti, lo, hi)
    shard = sharded.ItemShard(d(V[lo:hi]), lo, d(lp), d(li))
    dsp, dsi = d(sp), d(si)
    ok, ties_total = True, 0
    full = None
    if True:                                             # reference result: the unsharded evaluator
        full = ops.eval_mf(d(U), d(V), d(np.arange(nu, dtype=np.int32)), d(tp), d(ti), dsp, dsi, metric, K, return_ranks=True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for user_lo in range(0, nu, B):
        a, b = sharded.local_slice(B, rank, ws)          # the rows of the batch this rank owns
        mine = d(U[user_lo + a:user_lo + b])
        res, ranks, ties = sharded.evaluate_item_sharded(mine, shard, dsp, dsi, user_lo, B, metric, K, return_ranks=True)
        ok &= bool(torch.equal(ranks, full[1][user_lo:user_lo + B])) and bool(torch.equal(res, full[0][user_lo:user_lo + B]))
        ties_total += int(ties.item())
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("item-sharded evaluator, world %d: %d users x %d items (d=%d) in batches of %d: rows %s the unsharded "
              "evaluator's on every rank, %d tie rows, %.0f users/s (incl. per-batch checks)" % (
                  ws, nu, ni, dim, B, "bit-identical to" if int(flag) else "DIFFER from", ties_total, nu / dt), flush=True)
    assert int(flag) == 1
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
