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
    ok, ties_total, diff_rows = True, 0, 0
    full = None
    if True:                                             # reference result: the unsharded evaluator
        full = ops.eval_mf(d(U), d(V), d(np.arange(nu, dtype=np.int32)), d(tp), d(ti), dsp, dsi, metric, K, return_ranks=True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for user_lo in range(0, nu, B):
        a, b = sharded.local_slice(B, rank, ws)          # the rows of the batch this rank owns
        mine = d(U[user_lo + a:user_lo + b])
        res, ranks, ties = sharded.evaluate_item_sharded(mine, shard, dsp, dsi, user_lo, B, metric, K, return_ranks=True)
        # Rows with an exact fp32 score tie inside the global top K+1 are counted by the merge and ranked
        # score-descending / id-ascending; the unsharded evaluator's order among EQUAL scores is an artefact of
        # the reference's heap (evaluate.h:23-50), so such a row may legitimately differ -- but only by a
        # permutation / swap of items whose scores are equal.  Every other row must be bit-identical.
        want_r, want_m = full[1][user_lo:user_lo + B], full[0][user_lo:user_lo + B]
        bad = ((ranks != want_r).any(dim=1) | (res != want_m).any(dim=1)).nonzero().flatten()
        n_ties = int(ties.item())
        ties_total += n_ties
        diff_rows += int(bad.numel())
        if bad.numel() > n_ties:
            ok = False
        for b in bad.tolist():
            urow = torch.from_numpy(U[user_lo + b]).cuda().double()
            sa = (torch.from_numpy(V[ranks[b].cpu().numpy()]).cuda().double() @ urow)
            sb = (torch.from_numpy(V[want_r[b].cpu().numpy()]).cuda().double() @ urow)
            if float((sa - sb).abs().max()) > 1e-6:        # the two rankings carry the same score sequence
                ok = False
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("item-sharded evaluator, world %d: %d users x %d items (d=%d) in batches of %d: %s, %d tie rows "
              "(flagged by the merge), %d rows differ from the unsharded evaluator's (all of them equal-score "
              "permutations: %s), %.0f users/s (incl. per-batch checks)" % (
                  ws, nu, ni, dim, B, "OK" if int(flag) else "FAILED", ties_total, diff_rows, bool(int(flag)), nu / dt),
              flush=True)
    assert int(flag) == 1
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
