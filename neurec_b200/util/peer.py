"""Row-sharded tables across the GPUs of one NVLink domain (BASELINE config 5).

One process per GPU.  A table too large for one GPU is cut into equal row blocks; every rank
allocates its block and maps the blocks of the other ranks into its own address space, so that a
kernel can read and `RED` remote rows directly over NVLink (`nrc_mf_bpr_sgd_sharded`,
`nrc_mf_bpr_sgd_epoch`, the item-sharded evaluator).  `torch.distributed` only exchanges handles.

Two ways to get the mappings (``backend``):
  "ipc"   every shard is its own cudaMalloc (`nrc_shard_alloc`, never a slice of torch's caching
          allocator, so a handle is opened at most once per process) exported with CUDA IPC and
          opened by the peers with their own device current (`nrc_ipc_open`);
  "symm"  torch.distributed._symmetric_memory (CUDA VMM allocations exchanged as file descriptors),
          the plumbing NCCL-free collectives in PyTorch use.
"""
import ctypes
import os

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


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view memory it did not allocate."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def _view(ptr, shape, device):
    return torch.as_tensor(_RawCuda(ptr, shape), device=device)


class ShardSet:
    """One fp32 table cut into `world` row blocks.  `local` is this rank's block as a torch tensor
    (a view of the shard allocation), `ptrs[r]` the device address of rank r's block in THIS
    process (own memory for r == rank, a peer mapping otherwise)."""

    def __init__(self, local, ptrs, rank, backend, keep):
        self.local, self.ptrs, self.rank, self.backend, self._keep = local, list(ptrs), rank, backend, keep
        self.shape = tuple(local.shape)
        self.hot = self.hot_delta = None          # replicated head (enable_hot)
        self.n_hot = 0

    # ---- replicated head: rows [0, n_hot) of the GLOBAL table, the most popular items after relabel_by_degree ----
    def enable_hot(self, n_hot):
        """Replicate global rows [0, n_hot) on every rank (collective).  From here on the sharded step reads these
        rows from the replica and accumulates their deltas locally; call sync_hot() after every step and
        writeback_hot() before anything reads the owners' blocks (evaluation, checkpoint)."""
        rows, dim = self.shape
        n_hot = int(n_hot)
        if not 0 <= n_hot <= rows * self.world:
            raise ValueError("n_hot %d outside the table" % n_hot)
        if (n_hot * dim) % 4:
            raise ValueError("n_hot * dim must be a multiple of 4")
        self.n_hot = n_hot
        dev = self.local.device
        self.hot = torch.zeros((n_hot, dim), dtype=torch.float32, device=dev)
        self.hot_delta = torch.zeros((n_hot, dim), dtype=torch.float32, device=dev)
        lo, hi = self._own_hot_range()
        if hi > lo:
            self.hot[lo:hi] = self.local[lo - self.rank * rows:hi - self.rank * rows]
        if self.world > 1:
            dist.all_reduce(self.hot)                 # every row has exactly one owner: the sum IS the gather
        return self

    def _own_hot_range(self):
        rows = self.shape[0]
        lo = min(self.n_hot, self.rank * rows)
        return lo, min(self.n_hot, (self.rank + 1) * rows)

    def sync_hot(self):
        """End of a step: sum the ranks' deltas of the replicated rows (ONE all-reduce of n_hot * dim floats) and
        apply them to every replica (nrc_mf_hot_apply: hot += delta, delta = 0)."""
        if not self.n_hot:
            return
        if self.world > 1:
            dist.all_reduce(self.hot_delta)
        _lib.check(_lib.load().nrc_mf_hot_apply(ctypes.c_void_p(self.hot.data_ptr()), ctypes.c_void_p(self.hot_delta.data_ptr()),
                                                self.hot.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def writeback_hot(self):
        """Owners copy their replicated rows back into their blocks (the blocks' copies are stale in between)."""
        if not self.n_hot:
            return
        rows = self.shape[0]
        lo, hi = self._own_hot_range()
        if hi > lo:
            self.local[lo - self.rank * rows:hi - self.rank * rows] = self.hot[lo:hi]

    @property
    def world(self):
        return len(self.ptrs)

    def peer_view(self, r):
        """rank r's block as a tensor of this process (reads go over NVLink); debugging / tests."""
        return self.local if r == self.rank else _view(self.ptrs[r], self.shape, self.local.device)

    def ptr_array(self):
        return (ctypes.c_void_p * self.world)(*self.ptrs)

    def close(self):
        """Unmap the peers' blocks, then (after a barrier: nobody may still be writing) free ours."""
        if self._keep is None:
            return
        keep, self._keep = self._keep, None
        lib = _lib.load()
        if self.backend == "ipc":
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    lib.nrc_ipc_close(ctypes.c_void_p(p), 0)
            if dist.is_initialized():
                dist.barrier()
            self.local = None
            lib.nrc_shard_free(ctypes.c_void_p(keep))
        else:
            self.local = None


def alloc_sharded(rows, dim, backend=None):
    """Allocate this rank's [rows, dim] fp32 block (same shape on every rank) and map everyone
    else's.  Collective over the default process group."""
    # Default: symmetric memory (CUDA VMM allocations shared as file descriptors).  Measured on 2 x B200 with uniform ids
    # (tests/mgpu_peer_rate.py): random 512-byte rows of a peer block mapped that way run at NVLink rates, the same
    # block imported through legacy CUDA IPC (cudaIpcOpenMemHandle) at ~1/3 of that.
    backend = backend or os.environ.get("NRC_PEER_BACKEND", "symm")
    ws, rank = dist.get_world_size(), dist.get_rank()
    device = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    if backend == "ipc":
        handle = (ctypes.c_ubyte * 64)()
        ptr = ctypes.c_void_p()
        _lib.check(lib.nrc_shard_alloc(int(rows) * int(dim) * 4, ctypes.byref(ptr), handle))
        gathered = [None] * ws
        dist.all_gather_object(gathered, (bytes(handle), int(device.index)))
        ptrs = []
        for r, (hbytes, dev_index) in enumerate(gathered):
            if r == rank:
                ptrs.append(int(ptr.value))
                continue
            _lib.check(lib.nrc_enable_peer_access(dev_index))
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(hbytes)
            p = ctypes.c_void_p()
            _lib.check(lib.nrc_ipc_open(buf, 0, ctypes.byref(p)))
            ptrs.append(int(p.value))
        local = _view(ptr.value, (rows, dim), device)
        dist.barrier()
        return ShardSet(local, ptrs, rank, "ipc", int(ptr.value))
    if backend == "symm":
        import torch.distributed._symmetric_memory as symm_mem
        local = symm_mem.empty((rows, dim), dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(local, dist.group.WORLD.group_name)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        dist.barrier()
        return ShardSet(local, ptrs, rank, "symm", hdl)
    raise ValueError("unknown peer backend %r" % backend)


def single(local):
    """world = 1: the ShardSet of an ordinary tensor (no mapping)."""
    return ShardSet(local, [local.data_ptr()], 0, "local", None)


def relabel_by_degree(indices, num_items, n_hot=None):
    """Load-time item relabelling for the replicated head: the `n_hot` items of highest train degree get the new
    ids [0, n_hot) in descending degree (ties by old id); the others keep their relative order behind them (their
    order is irrelevant to the head -- and a popularity-sorted tail would put all warm rows next to each other in
    memory).  n_hot=None sorts the whole catalogue.  `indices` are the train CSR's item ids; returns
    (new_id_of_old int32 [num_items], degree of every NEW id int64 [num_items]).  The reference remaps raw ids to
    dense ones at load time as well (data/dataset.py:88-110); this only fixes the order of that remap."""
    indices = np.asarray(indices)
    num_items = int(num_items)
    deg = np.bincount(indices, minlength=num_items).astype(np.int64)
    order = np.argsort(-deg, kind="stable")
    if n_hot is not None and n_hot < num_items:
        head = order[:int(n_hot)]
        rest = np.ones(num_items, bool)
        rest[head] = False
        order = np.concatenate([head, np.nonzero(rest)[0]])
    new_of_old = np.empty(num_items, np.int32)
    new_of_old[order] = np.arange(num_items, dtype=np.int32)
    return new_of_old, deg[order]


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
