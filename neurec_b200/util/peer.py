"""Row-sharded tables across the GPUs of one NVLink domain (BASELINE config 5).

One process per GPU.  A table too large for one GPU is cut into equal row blocks; every rank
allocates its block and maps the blocks of the other ranks into its address space through CUDA
IPC (PyTorch's storage-sharing plumbing), so that a kernel can read and `RED` remote rows directly
over NVLink (`nrc_mf_bpr_sgd_sharded`).  `torch.distributed` is only used to exchange the handles.
"""
import numpy as np
import torch
import torch.distributed as dist

from .. import _lib


def rows_per_shard(total_rows, world):
    """Equal blocks; the last shard is padded (ids >= total_rows are never generated)."""
    return (int(total_rows) + world - 1) // world


def owner_of(ids, per_shard):
    """Owner rank of global row ids."""
    return np.asarray(ids) // int(per_shard)


def open_peer_shards(local):
    """All-gather CUDA IPC handles of `local` (this rank's [rows_per_shard, dim] block, the same
    shape on every rank) and return one tensor per rank: `local` itself at this rank's position,
    peer mappings elsewhere.  Peer access is enabled for kernels of the current device.  Keep the
    returned tensors alive as long as kernels may touch them."""
    from torch.multiprocessing.reductions import reduce_tensor
    ws, rank = dist.get_world_size(), dist.get_rank()
    assert local.is_cuda and local.is_contiguous()
    gathered = [None] * ws
    dist.all_gather_object(gathered, reduce_tensor(local))     # (rebuild function, IPC handle + layout)
    out = []
    for r, (rebuild, args) in enumerate(gathered):
        if r == rank:
            out.append(local)
            continue
        t = rebuild(*args)                                       # cudaIpcOpenMemHandle on the owner's device
        assert t.shape == local.shape and t.dtype == local.dtype
        _lib.check(_lib.load().nrc_enable_peer_access(int(t.device.index)))
        out.append(t)
    dist.barrier()
    return out


def route_triplets_to_user_owner(users, pos, neg, users_per_shard):
    """Exchange step for a GLOBAL triplet stream ("partition triplets by user
    owner"): every rank passes the triplets it produced, every rank receives the triplets whose user
    it owns, in (source rank, original order).  numpy int32 arrays in and out; works on any backend
    (gloo on CPU, NCCL via object collectives)."""
    ws, rank = dist.get_world_size(), dist.get_rank()
    users, pos, neg = (np.asarray(a, dtype=np.int32) for a in (users, pos, neg))
    own = owner_of(users, users_per_shard)
    if own.size and (own.min() < 0 or own.max() >= ws):
        raise ValueError("user id outside the sharded table")
    parts = [np.stack([users[own == r], pos[own == r], neg[own == r]]) for r in range(ws)]
    gathered = [None] * ws
    dist.all_gather_object(gathered, parts)        # small host-side metadata path; ids only
    mine = [gathered[src][rank] for src in range(ws)]
    cat = np.concatenate(mine, axis=1) if mine else np.zeros((3, 0), np.int32)
    return cat[0].copy(), cat[1].copy(), cat[2].copy()
