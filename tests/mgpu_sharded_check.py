"""Run under torchrun with N >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29521 tests/mgpu_sharded_check.py

Row-sharded BPR + SGD step (nrc_mf_bpr_sgd_sharded, BASELINE config 5): every rank owns one row
block of the user and of the item table, maps the other blocks through CUDA IPC and trains its own
users' triplets; item rows of other ranks are read and RED-updated over NVLink by the same kernel.
The triplets are built so that no row repeats anywhere in the step, which makes the in-place step
exactly the textbook one; the gathered tables must match the numpy restatement (oracle/tf_math.py)
applied to the full tables."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = dist.get_rank(), dist.get_world_size()
    from neurec_b200 import ops
    from neurec_b200.util import peer
    from oracle import tf_math

    # stage 0: plumbing only -- peer shards are mapped and passed, but every triplet touches local rows
    nu_l, ni_l, dim = 1000, 1000, 64
    g = torch.Generator(device="cuda").manual_seed(rank)
    myU = torch.randn(nu_l, dim, device="cuda", generator=g) * 0.1
    myV = torch.randn(ni_l, dim, device="cuda", generator=g) * 0.1
    Us, Vs = peer.open_peer_shards(myU), peer.open_peer_shards(myV)
    if rank == 0:
        print("peer shards mapped:", [type(t).__name__ for t in Us], flush=True)
    loc = lambda n, per: (torch.randperm(per, device="cuda")[:n] + rank * per).to(torch.int32)
    loss = torch.zeros(1, device="cuda")
    before = myV.clone()
    ops.mf_bpr_sgd_sharded(Us, Vs, rank, loc(200, nu_l), loc(200, ni_l), loc(200, ni_l), 0.05, 0.0, loss)
    torch.cuda.synchronize()
    if rank == 0:
        print("stage 0 (local rows through the sharded entry point): ok, table moved by %.2e" %
              float((myV - before).abs().max()), flush=True)
    dist.barrier()
    # stage 1: remote rows only (items of the NEXT rank)
    nxt = (rank + 1) % ws
    rem = lambda n: (torch.randperm(ni_l, device="cuda")[:n] + nxt * ni_l).to(torch.int32)
    ops.mf_bpr_sgd_sharded(Us, Vs, rank, loc(200, nu_l), rem(200), rem(200), 0.05, 0.0, loss)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("stage 1 (item rows of the next rank: peer loads + peer REDs over NVLink): ok, my table moved by %.2e"
              % float((myV - before).abs().max()), flush=True)
    del Us, Vs
    dist.barrier()

    for dim in (128, 64):
        nu_l, ni_l, per_rank = 3000, 5000, 1200          # rows per shard, triplets per rank
        nu, ni = nu_l * ws, ni_l * ws
        rs = np.random.RandomState(42)                    # same stream on every rank: full tables + full triplet set
        U = (rs.randn(nu, dim) * 0.1).astype(np.float32)
        V = (rs.randn(ni, dim) * 0.1).astype(np.float32)
        items = rs.permutation(ni)[:2 * per_rank * ws].astype(np.int32)      # every item row at most once
        users = np.concatenate([r * nu_l + rs.permutation(nu_l)[:per_rank] for r in range(ws)]).astype(np.int32)
        pos, neg = items[:per_rank * ws], items[per_rank * ws:]
        lr, reg = 0.05, 0.01

        # oracle on the full tables (no repeated row => order-free)
        pu, qi, qj = U[users], V[pos], V[neg]
        x = (pu * qi).sum(1) - (pu * qj).sum(1)
        want_loss, g = tf_math.pairwise_loss_and_grad("bpr", x)
        g = g[:, None].astype(np.float32)
        Uw, Vw = U.copy(), V.copy()
        Uw[users] -= np.float32(lr) * (g * (qi - qj) + np.float32(reg) * pu)
        Vw[pos] -= np.float32(lr) * (g * pu + np.float32(reg) * qi)
        Vw[neg] -= np.float32(lr) * (-g * pu + np.float32(reg) * qj)

        myU = torch.from_numpy(U[rank * nu_l:(rank + 1) * nu_l].copy()).cuda()
        myV = torch.from_numpy(V[rank * ni_l:(rank + 1) * ni_l].copy()).cuda()
        Us, Vs = peer.open_peer_shards(myU), peer.open_peer_shards(myV)
        sl = slice(rank * per_rank, (rank + 1) * per_rank)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        loss = torch.zeros(1, device="cuda")
        ops.mf_bpr_sgd_sharded(Us, Vs, rank, d(users[sl]), d(pos[sl]), d(neg[sl]), lr, reg, loss)
        torch.cuda.synchronize()
        dist.barrier()                                    # every rank's remote REDs have landed
        gU = [torch.empty_like(myU) for _ in range(ws)]
        gV = [torch.empty_like(myV) for _ in range(ws)]
        dist.all_gather(gU, myU)
        dist.all_gather(gV, myV)
        gotU, gotV = torch.cat(gU).cpu().numpy(), torch.cat(gV).cpu().numpy()
        dist.all_reduce(loss)
        remote = float(np.mean(peer.owner_of(np.concatenate([pos[sl], neg[sl]]), ni_l) != rank))
        okU = float(np.abs(gotU - Uw).max()); okV = float(np.abs(gotV - Vw).max())
        # reg term of the loss: 0.5*reg*(|pu|^2+|qi|^2+|qj|^2) summed (tool.py:216-217)
        want_total = float(np.sum(want_loss, dtype=np.float64) + 0.5 * reg * np.sum(pu * pu + qi * qi + qj * qj, dtype=np.float64))
        ok = okU < 2e-6 and okV < 2e-6 and abs(float(loss) - want_total) < 1e-3 * abs(want_total)
        moved = float(np.abs(Vw - V).max())
        if rank == 0:
            print("dim %d world %d: max|dU| %.2e max|dV| %.2e (tables moved by %.2e), loss %.4f vs %.4f, "
                  "%.0f%% of this rank's item rows are remote -> %s" % (dim, ws, okU, okV, moved, float(loss),
                                                                      want_total, 100 * remote, "OK" if ok else "MISMATCH"))
        assert ok, (okU, okV, float(loss), want_total)
        del Us, Vs
        dist.barrier()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
