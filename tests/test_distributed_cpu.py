"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous user sharding and the
order-preserving all-gather of per-user rows used by the sharded evaluator."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurec_b200.evaluator import sharded


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 0.5
    a, b = sharded.local_slice(n, rank, ws)
    got = sharded.gather_rows(full[a:b].clone(), n)
    s, cnt = sharded.reduce_sums(full[a:b].clone())
    ok = torch.equal(got, full) and cnt == n and torch.allclose(s, full.double().sum(0))
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_local_slice_partitions():
    for n in (0, 1, 7, 943, 29858):
        for ws in (1, 2, 3, 8):
            cuts = [sharded.local_slice(n, r, ws) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_gather_rows_world_size_2_gloo():
    ws, port = 2, _free_port()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(ws, port, 943, out), nprocs=ws, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_single_process_is_identity():
    x = torch.rand(5, 4)
    assert sharded.gather_rows(x, 5) is x
    s, n = sharded.reduce_sums(x)
    assert n == 5 and np.allclose(s.numpy(), x.double().sum(0).numpy())


def _route_worker(rank, ws, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from neurec_b200.util import peer
    per = peer.rows_per_shard(1001, ws)                       # 501 rows per shard
    rs = np.random.RandomState(10 + rank)
    n = 300 + 17 * rank
    users = rs.randint(0, 1001, n).astype(np.int32)
    pos = (users * 3 + rank).astype(np.int32)                 # payload tied to the user id and the source
    neg = (users * 5 + 7).astype(np.int32)
    u, p, q = peer.route_triplets_to_user_owner(users, pos, neg, per)
    ok = bool(np.all(peer.owner_of(u, per) == rank))          # only my users arrive
    ok &= bool(np.all(q == u * 5 + 7)) and bool(np.all((p - u * 3 >= 0) & (p - u * 3 < ws)))
    counts = [None] * ws
    dist.all_gather_object(counts, (n, len(u), int(np.sum(peer.owner_of(users, per) == rank))))
    ok &= sum(c[0] for c in counts) == sum(c[1] for c in counts)       # nothing lost or duplicated
    # triplets keep (source rank, original order): my own part is a subsequence of what I sent
    mine_sent = users[peer.owner_of(users, per) == rank]
    src = p - u * 3
    ok &= bool(np.array_equal(u[src == rank], mine_sent))
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_route_triplets_to_user_owner_world_size_2_gloo():
    """Host logic of the row-sharded training path (BASELINE config 5): triplets are exchanged so
    that every rank trains the users whose rows it owns."""
    ws = 2
    out = mp.Manager().dict()
    mp.spawn(_route_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    assert [out[r] for r in range(ws)] == [1, 1]


def test_item_shard_csr_restriction_partitions_every_row():
    """Host side of the item-sharded evaluator (evaluator/sharded.py::ItemShard.restrict_csr): the train CSR
    cut by item range keeps every (user, item) entry in exactly one shard, with local ids, rows still sorted."""
    import numpy as np
    from neurec_b200.evaluator import sharded
    rs = np.random.RandomState(0)
    nu, ni, G = 50, 1000, 8
    rows = [np.unique(rs.randint(0, ni, rs.randint(0, 40))) for _ in range(nu)]
    ptr = np.zeros(nu + 1, np.int64); ptr[1:] = np.cumsum([len(r) for r in rows])
    idx = np.concatenate(rows).astype(np.int32)
    per = (ni + G - 1) // G
    rebuilt = [[] for _ in range(nu)]
    for g in range(G):
        lo, hi = g * per, min(ni, (g + 1) * per)
        lp, li = sharded.ItemShard.restrict_csr(ptr, idx, lo, hi)
        assert lp[0] == 0 and len(lp) == nu + 1
        for u in range(nu):
            loc = li[lp[u]:lp[u + 1]]
            assert (np.diff(loc) > 0).all() and (loc >= 0).all() and (loc < hi - lo).all()
            rebuilt[u].extend((loc + lo).tolist())
    for u in range(nu):
        assert rebuilt[u] == rows[u].tolist()
    # contiguous user blocks used by the all-gather of user rows
    blocks = [sharded.local_slice(37, r, 4) for r in range(4)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 37 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_relabel_by_degree_orders_items_by_popularity():
    """Load-time relabelling behind the replicated head: descending train degree, ties by old id."""
    from neurec_b200.util import peer
    idx = np.array([5, 5, 5, 2, 2, 7, 0, 0, 9], np.int32)
    new_of_old, deg = peer.relabel_by_degree(idx, 10)
    assert new_of_old[5] == 0 and new_of_old[0] == 1 and new_of_old[2] == 2          # 3, 2 (id 0 first), 2
    assert new_of_old[7] == 3 and new_of_old[9] == 4
    assert sorted(new_of_old.tolist()) == list(range(10))
    assert deg.tolist() == [3, 2, 2, 1, 1, 0, 0, 0, 0, 0]
    assert np.all(np.diff(np.bincount(new_of_old[idx], minlength=10)) <= 0)
    head, deg2 = peer.relabel_by_degree(idx, 10, n_hot=2)           # only the head is sorted, the tail keeps its order
    assert head[5] == 0 and head[0] == 1
    assert [int(np.nonzero(head == k)[0][0]) for k in range(2, 10)] == [1, 2, 3, 4, 6, 7, 8, 9]
    assert deg2[:2].tolist() == [3, 2]
