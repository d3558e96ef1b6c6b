"""Row-sharded tables across the GPUs of one NVLink domain (BASELINE config 5).

One process per GPU.  A table too large for one GPU is cut into equal row blocks; every rank
allocates its block and maps the blocks of the other ranks into its address space through CUDA
IPC (PyTorch's storage-sharing plumbing), so that a kernel can read and `RED` remote rows directly
over NVLink (`nrc_mf_bpr_sgd_sharded`).  `torch.distributed` is only used to exchange the handles.
"""
import ctypes

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


class PeerShard:
    """A row block that lives on another rank's GPU, mapped into this process with CUDA IPC
    (nrc_ipc_open, importing device current).  Only what the kernels need: data_ptr() and shape."""

    def __init__(self, ptr, offset, shape, dtype):
        self._ptr, self._offset, self.shape, self.dtype = ptr, offset, tuple(shape), dtype

    def data_ptr(self):
        return self._ptr

    def close(self):
        if self._ptr:
            _lib.load().nrc_ipc_close(ctypes.c_void_p(self._ptr), self._offset)
            self._ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open_peer_shards(local):
    """All-gather CUDA IPC handles of `local` (this rank's [rows_per_shard, dim] block, the same
    shape on every rank) and return one entry per rank: `local` itself at this rank's position,
    PeerShard mappings elsewhere (opened with THIS rank's device current, so its kernels can read
    and RED the rows over NVLink).  Keep `local` alive on its owner while any peer uses it.

    Round-1 status: the first version went through torch's storage sharing, which opens the handle
    with the owner's device current; kernels then faulted on the mapping.  This version has not run
    on a multi-GPU box yet (tests/mgpu_sharded_check.py is the check)."""
    ws, rank = dist.get_world_size(), dist.get_rank()
    assert local.is_cuda and local.is_contiguous()
    lib = _lib.load()
    handle = (ctypes.c_ubyte * 64)()
    offset = ctypes.c_int64(0)
    _lib.check(lib.nrc_ipc_export(ctypes.c_void_p(local.data_ptr()), handle, ctypes.byref(offset)))
    info = (bytes(handle), int(offset.value), tuple(local.shape), str(local.dtype), int(local.device.index))
    gathered = [None] * ws
    dist.all_gather_object(gathered, info)
    out = []
    for r, (hbytes, off, shape, dtype, dev_index) in enumerate(gathered):
        if r == rank:
            out.append(local)
            continue
        assert tuple(shape) == tuple(local.shape) and dtype == str(local.dtype)
        _lib.check(lib.nrc_enable_peer_access(dev_index))
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(hbytes)
        ptr = ctypes.c_void_p()
        _lib.check(lib.nrc_ipc_open(buf, off, ctypes.byref(ptr)))
        out.append(PeerShard(int(ptr.value), off, shape, local.dtype))
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
