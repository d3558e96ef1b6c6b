"""Run under torchrun with N >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29521 tests/mgpu_sharded_check.py [ipc|symm]

Row-sharded BPR + SGD (nrc_mf_bpr_sgd_sharded / nrc_mf_bpr_sgd_epoch, BASELINE config 5): every
rank owns one row block of the user and of the item table, maps the other blocks (CUDA IPC or
symmetric memory) and trains its own users' triplets; item rows of other ranks are read and
RED-updated over NVLink by the same kernel.
  stage 0  peer blocks readable through plain tensor views (mapping works at all)
  stage 1  local rows through the sharded entry point
  stage 2  remote rows only (peer loads + peer REDs)
  stage 3  no-duplicate batches: the gathered tables equal the numpy restatement on the FULL tables
  stage 5  stage 4 again with a replicated head (ShardSet.enable_hot / sync_hot / writeback_hot)
  stage 4  the CSR-fed kernel (sampler + shuffle fused in) on sharded tables equals the same kernel
           run by ONE rank on the full item table with every rank's epoch applied in rank order
           (only checked on duplicate-free epochs where the order does not matter)
Every stage prints its own line so a fault names the stage it happened in."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def say(rank, *a):
    if rank == 0:
        print(*a, flush=True)


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("NRC_PEER_BACKEND", "ipc")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = dist.get_rank(), dist.get_world_size()
    from neurec_b200 import ops
    from neurec_b200.util import peer
    from oracle import tf_math
    import oracle
    say(rank, "backend:", backend, "world:", ws)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    # ---- stage 0: mapping
    nu_l, ni_l, dim = 1000, 1000, 64
    US, VS = peer.alloc_sharded(nu_l, dim, backend), peer.alloc_sharded(ni_l, dim, backend)
    g = torch.Generator(device="cuda").manual_seed(rank)
    US.local.copy_(torch.randn(nu_l, dim, device="cuda", generator=g) * 0.1)
    VS.local.copy_(torch.randn(ni_l, dim, device="cuda", generator=g) * 0.1 + rank)
    torch.cuda.synchronize(); dist.barrier()
    nxt = (rank + 1) % ws
    seen = float(VS.peer_view(nxt).mean())
    torch.cuda.synchronize()
    ok0 = abs(seen - nxt) < 0.05
    say(rank, "stage 0 (peer block readable through a tensor view): mean %.3f, expected ~%d -> %s" %
        (seen, nxt, "ok" if ok0 else "MISMATCH"))
    assert ok0
    dist.barrier()

    # ---- stage 1: local rows through the sharded entry point
    loc = lambda n, per: (torch.randperm(per, device="cuda")[:n] + rank * per).to(torch.int32)
    loss = torch.zeros(1, device="cuda")
    before = VS.local.clone()
    ops.mf_bpr_sgd_sharded(US, VS, rank, loc(200, nu_l), loc(200, ni_l), loc(200, ni_l), 0.05, 0.0, loss)
    torch.cuda.synchronize()
    say(rank, "stage 1 (local rows through the sharded entry point): ok, table moved by %.2e" %
        float((VS.local - before).abs().max()))
    dist.barrier()

    # ---- stage 2: remote rows only
    before = VS.local.clone()
    rem = lambda n: (torch.randperm(ni_l, device="cuda")[:n] + nxt * ni_l).to(torch.int32)
    ops.mf_bpr_sgd_sharded(US, VS, rank, loc(200, nu_l), rem(200), rem(200), 0.05, 0.0, loss)
    torch.cuda.synchronize()
    dist.barrier()
    moved = float((VS.local - before).abs().max())
    say(rank, "stage 2 (item rows of the next rank: peer loads + peer REDs over NVLink): my block moved by %.2e -> %s"
        % (moved, "ok" if moved > 0 else "NOT UPDATED"))
    assert moved > 0
    dist.barrier()
    US.close(); VS.close()

    # ---- stage 3: full no-duplicate step vs numpy on the full tables
    for dim in (128, 64):
        nu_l, ni_l, per_rank = 3000, 5000, 1200          # rows per shard, triplets per rank
        nu, ni = nu_l * ws, ni_l * ws
        rs = np.random.RandomState(42)                    # same stream on every rank
        U = (rs.randn(nu, dim) * 0.1).astype(np.float32)
        V = (rs.randn(ni, dim) * 0.1).astype(np.float32)
        items = rs.permutation(ni)[:2 * per_rank * ws].astype(np.int32)      # every item row at most once
        users = np.concatenate([r * nu_l + rs.permutation(nu_l)[:per_rank] for r in range(ws)]).astype(np.int32)
        pos, neg = items[:per_rank * ws], items[per_rank * ws:]
        lr, reg = 0.05, 0.01
        pu, qi, qj = U[users], V[pos], V[neg]
        x = (pu * qi).sum(1) - (pu * qj).sum(1)
        want_loss, gg = tf_math.pairwise_loss_and_grad("bpr", x)
        gg = gg[:, None].astype(np.float32)
        Uw, Vw = U.copy(), V.copy()
        Uw[users] -= np.float32(lr) * (gg * (qi - qj) + np.float32(reg) * pu)
        Vw[pos] -= np.float32(lr) * (gg * pu + np.float32(reg) * qi)
        Vw[neg] -= np.float32(lr) * (-gg * pu + np.float32(reg) * qj)

        US, VS = peer.alloc_sharded(nu_l, dim, backend), peer.alloc_sharded(ni_l, dim, backend)
        US.local.copy_(d(U[rank * nu_l:(rank + 1) * nu_l])); VS.local.copy_(d(V[rank * ni_l:(rank + 1) * ni_l]))
        torch.cuda.synchronize(); dist.barrier()
        sl = slice(rank * per_rank, (rank + 1) * per_rank)
        loss = torch.zeros(1, device="cuda")
        ops.mf_bpr_sgd_sharded(US, VS, rank, d(users[sl]), d(pos[sl]), d(neg[sl]), lr, reg, loss)
        torch.cuda.synchronize()
        dist.barrier()                                    # every rank's remote REDs have landed
        gU = [torch.empty_like(US.local) for _ in range(ws)]
        gV = [torch.empty_like(VS.local) for _ in range(ws)]
        dist.all_gather(gU, US.local.contiguous())
        dist.all_gather(gV, VS.local.contiguous())
        gotU, gotV = torch.cat(gU).cpu().numpy(), torch.cat(gV).cpu().numpy()
        dist.all_reduce(loss)
        remote = float(np.mean(peer.owner_of(np.concatenate([pos[sl], neg[sl]]), ni_l) != rank))
        okU = float(np.abs(gotU - Uw).max()); okV = float(np.abs(gotV - Vw).max())
        want_total = float(np.sum(want_loss, dtype=np.float64) + 0.5 * reg * np.sum(pu * pu + qi * qi + qj * qj, dtype=np.float64))
        ok = okU < 2e-6 and okV < 2e-6 and abs(float(loss) - want_total) < 1e-3 * abs(want_total)
        say(rank, "stage 3 dim %d world %d: max|dU| %.2e max|dV| %.2e (tables moved by %.2e), loss %.4f vs %.4f, "
            "%.0f%% of this rank's item rows are remote -> %s" % (dim, ws, okU, okV, float(np.abs(Vw - V).max()),
                                                                float(loss), want_total, 100 * remote, "OK" if ok else "MISMATCH"))
        assert ok, (okU, okV, float(loss), want_total)

        # ---- stage 4: CSR-fed kernel.  Each local user has ONE positive; positives and the drawn
        # negatives are checked duplicate-free across ranks on the host first (else the stage is skipped).
        rs = np.random.RandomState(7 + dim)
        n_loc = 40
        tp = np.arange(n_loc + 1, dtype=np.int64)                       # local users 0..n_loc-1, one item each
        all_pos = rs.permutation(ni)[:n_loc * ws].astype(np.int32)
        tis = [all_pos[r * n_loc:(r + 1) * n_loc] for r in range(ws)]
        pus = np.arange(n_loc, dtype=np.int32)
        for seed0 in range(100, 400, ws):                              # first seed set with a duplicate-free epoch
            epochs = [oracle.epoch_build(tp, tis[r], pus, tis[r], 1, ni, True, True, seed0 + r, 3) for r in range(ws)]
            flat = np.concatenate([np.concatenate([e[1], e[2][:, 0]]) for e in epochs])
            dup_free = len(np.unique(flat)) == len(flat)
            if dup_free:
                break
        U2 = (rs.randn(nu, dim) * 0.1).astype(np.float32)
        V2 = (rs.randn(ni, dim) * 0.1).astype(np.float32)
        US.local.copy_(d(U2[rank * nu_l:(rank + 1) * nu_l])); VS.local.copy_(d(V2[rank * ni_l:(rank + 1) * ni_l]))
        torch.cuda.synchronize(); dist.barrier()
        loss = torch.zeros(1, device="cuda")
        ops.mf_bpr_sgd_epoch(US.local, VS, d(tp), d(tis[rank]), d(pus), d(tis[rank]), ni, True, seed0 + rank, 3, 0, n_loc,
                             lr, reg, loss)
        torch.cuda.synchronize(); dist.barrier()
        dist.all_gather(gU, US.local.contiguous()); dist.all_gather(gV, VS.local.contiguous())
        gotU, gotV = torch.cat(gU).cpu().numpy(), torch.cat(gV).cpu().numpy()
        Uw, Vw = U2.copy(), V2.copy()
        for r in range(ws):                                            # numpy restatement, rank by rank
            eu, ei, ej = epochs[r][0].astype(np.int64) + r * nu_l, epochs[r][1], epochs[r][2][:, 0]
            pu, qi, qj = Uw[eu], Vw[ei], Vw[ej]
            x = (pu * qi).sum(1) - (pu * qj).sum(1)
            _, gg = tf_math.pairwise_loss_and_grad("bpr", x)
            gg = gg[:, None].astype(np.float32)
            Uw[eu] -= np.float32(lr) * (gg * (qi - qj) + np.float32(reg) * pu)
            Vw[ei] -= np.float32(lr) * (gg * pu + np.float32(reg) * qi)
            Vw[ej] -= np.float32(lr) * (-gg * pu + np.float32(reg) * qj)
        okU = float(np.abs(gotU - Uw).max()); okV = float(np.abs(gotV - Vw).max())
        ok = (okU < 2e-6 and okV < 2e-6) or not dup_free
        say(rank, "stage 4 dim %d (sampler + shuffle fused, CSR-fed, sharded): dup-free %s, max|dU| %.2e max|dV| %.2e -> %s"
            % (dim, dup_free, okU, okV, "OK" if ok else "MISMATCH"))
        assert ok
        # ---- stage 5: the same epoch with a REPLICATED HEAD (half of the item table): reads from the replica,
        # deltas summed over the ranks by sync_hot's all-reduce, owners write back -> the same tables
        US.local.copy_(d(U2[rank * nu_l:(rank + 1) * nu_l])); VS.local.copy_(d(V2[rank * ni_l:(rank + 1) * ni_l]))
        torch.cuda.synchronize(); dist.barrier()
        VS.enable_hot(ni // 2)
        loss = torch.zeros(1, device="cuda")
        ops.mf_bpr_sgd_epoch(US.local, VS, d(tp), d(tis[rank]), d(pus), d(tis[rank]), ni, True, seed0 + rank, 3, 0, n_loc,
                             lr, reg, loss)
        VS.sync_hot(); VS.writeback_hot()
        torch.cuda.synchronize(); dist.barrier()
        dist.all_gather(gU, US.local.contiguous()); dist.all_gather(gV, VS.local.contiguous())
        gotU, gotV = torch.cat(gU).cpu().numpy(), torch.cat(gV).cpu().numpy()
        okU = float(np.abs(gotU - Uw).max()); okV = float(np.abs(gotV - Vw).max())
        ok = (okU < 2e-6 and okV < 2e-6) or not dup_free
        say(rank, "stage 5 dim %d (replicated head of %d rows, all-reduced deltas): max|dU| %.2e max|dV| %.2e -> %s"
            % (dim, ni // 2, okU, okV, "OK" if ok else "MISMATCH"))
        assert ok
        VS.n_hot = 0
        US.close(); VS.close()
        dist.barrier()
    say(rank, "ALL STAGES OK (%s, world %d)" % (backend, ws))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
