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
