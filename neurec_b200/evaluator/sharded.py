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
