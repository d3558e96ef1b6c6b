"""Multi-GPU evaluation: users are independent, so the test users are split into contiguous
blocks, one per rank (tables replicated), and the per-user metric rows are all-gathered back in
the original user order before the fp32 mean -- the result string is therefore bit-identical for
any number of GPUs.  (For catalogues where gathering [num_users, M*K] rows is too much, use
``reduce_sums``: fp64 partial sums + one all-reduce.)  One process per GPU, torch.distributed
(NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_slice(n, rank, world_size):
    """Contiguous block of rank `rank` out of n items (block sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local_rows, n_total):
    """All-gather per-user rows [n_local, C] -> [n_total, C] in rank (= user) order."""
    rank, ws = world()
    if ws == 1:
        return local_rows
    cols = local_rows.shape[1]
    width = (n_total + ws - 1) // ws
    pad = torch.zeros((width, cols), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    parts = []
    for r in range(ws):
        a, b = local_slice(n_total, r, ws)
        parts.append(out[r][:b - a])
    return torch.cat(parts, dim=0)


def reduce_sums(local_rows):
    """Scalable alternative: (sum over all users in fp64 [C], total user count)."""
    s = local_rows.to(torch.float64).sum(dim=0)
    n = torch.tensor([float(local_rows.shape[0])], dtype=torch.float64, device=local_rows.device)
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(s)
        dist.all_reduce(n)
    return s, int(n.item())


# ----------------------------------------------------------------------------------------
# Item-sharded evaluation: tables that exceed one GPU (BASELINE configs[4] scale)
# ----------------------------------------------------------------------------------------
class ItemShard:
    """What one rank holds: rows [item_lo, item_lo + n_local) of the item table and the train CSR
    restricted to that item range with LOCAL item ids (rows = global user ids), both resident."""

    def __init__(self, item_table, item_lo, train_indptr, train_indices):
        self.V, self.lo = item_table, int(item_lo)
        self.train_ptr, self.train_idx = train_indptr, train_indices

    @staticmethod
    def restrict_csr(indptr, indices, lo, hi):
        """Host-side (load time, like the reference's dataset split): CSR of the entries in
        [lo, hi), shifted to local ids.  numpy in, numpy out."""
        import numpy as np
        indptr, indices = np.asarray(indptr, np.int64), np.asarray(indices, np.int64)
        keep = (indices >= lo) & (indices < hi)
        rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))[keep]
        ptr = np.zeros(len(indptr), np.int64)
        np.add.at(ptr, rows + 1, 1)
        idx = (indices[keep] - lo).astype(np.int32)
        if idx.size == 0:
            idx = np.zeros(1, np.int32)
        return np.cumsum(ptr), idx


def shard_candidates(user_rows, shard, user_lo, top_k):
    """This rank's contribution for a batch of CONSECUTIVE users [user_lo, user_lo + B): the K+1 best
    unmasked items of its shard per user as (global ids i32 [B, K+1], exact scores f32 [B, K+1])."""
    from .. import ops
    B = user_rows.shape[0]
    dev = user_rows.device
    rows = torch.arange(B, dtype=torch.int32, device=dev)
    tp = shard.train_ptr[user_lo:user_lo + B + 1]                      # a view: CSR rows of the batch
    k1 = min(top_k + 1, shard.V.shape[0])
    nothing = torch.zeros(B + 1, dtype=torch.int64, device=dev)       # no truth needed for ranks only
    if ops.use_tensor_core_eval(shard.V.shape[0], shard.V.shape[1], k1, B):      # large shard: tcgen05 candidate pass
        _, local = ops.eval_mf_tc(user_rows, shard.V, rows, tp, shard.train_idx, nothing, shard.train_idx,
                                  ["Precision"], k1, return_ranks=True)
    else:
        _, local = ops.eval_mf(user_rows, shard.V, rows, tp, shard.train_idx, nothing, shard.train_idx,
                               ["Precision"], k1, return_ranks=True, want_results=False)
    if k1 < top_k + 1:
        local = torch.cat([local, torch.full((B, top_k + 1 - k1), -1, dtype=torch.int32, device=dev)], 1).contiguous()
    scores = ops.mf_score_pairs(user_rows, shard.V, local, tp, shard.train_idx)
    gids = torch.where(local >= 0, local + shard.lo, local).contiguous()
    return gids, scores


def merge_and_score(all_ids, all_scores, test_indptr, test_indices, user_lo, metric, top_k, return_ranks=False):
    """Home-rank merge of the per-shard lists ([G][B, K+1] each) + metrics; the test CSR has GLOBAL
    item ids and rows = global user ids."""
    from .. import ops
    ids = torch.cat(all_ids, dim=1).contiguous()
    sc = torch.cat(all_scores, dim=1).contiguous()
    B = ids.shape[0]
    return ops.eval_merge_candidates(ids, sc, test_indptr[user_lo:user_lo + B + 1], test_indices, metric, top_k,
                                     return_ranks)


def evaluate_item_sharded(user_rows_local, shard, test_indptr, test_indices, user_lo, n_batch, metric, top_k,
                          return_ranks=False):
    """One evaluation batch with BOTH tables row-sharded over the ranks (collective over the default
    group): `user_rows_local` are this rank's rows of the batch's users (consecutive global ids
    [user_lo, user_lo + n_batch), owned by ranks in rank order).  ONE all-gather of the user rows
    [B, d], the local candidate pass, ONE all-gather of the [B, K+1] (id, score) lists, merge.
    Every rank returns the full batch's rows (identical on all ranks)."""
    rank, ws = world()
    if ws == 1:
        rows = user_rows_local
    else:
        d = user_rows_local.shape[1]
        counts = [local_slice(n_batch, r, ws) for r in range(ws)]
        width = max(b - a for a, b in counts)
        pad = torch.zeros((width, d), dtype=torch.float32, device=user_rows_local.device)
        pad[:user_rows_local.shape[0]] = user_rows_local
        out = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(out, pad)                                      # the batch's user rows: B * d * 4 bytes
        rows = torch.cat([out[r][:b - a] for r, (a, b) in enumerate(counts)], dim=0).contiguous()
    gids, scores = shard_candidates(rows, shard, user_lo, top_k)
    if ws == 1:
        all_ids, all_scores = [gids], [scores]
    else:
        all_ids = [torch.empty_like(gids) for _ in range(ws)]
        all_scores = [torch.empty_like(scores) for _ in range(ws)]
        dist.all_gather(all_ids, gids)                                 # [B, K+1] (id, score) per rank
        dist.all_gather(all_scores, scores)
    return merge_and_score(all_ids, all_scores, test_indptr, test_indices, user_lo, metric, top_k, return_ranks)
