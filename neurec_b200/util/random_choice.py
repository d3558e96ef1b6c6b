"""randint_choice / batch_randint_choice with the reference's signature and errors
(util/cython/random_choice.pyx:20-89), drawn by the sm_100a Philox sampler.

The reference's stream (glibc rand(), never seeded) cannot be reproduced by a parallel
generator; the contract is: uniform over [0, high) minus `exclusion`, with or without
replacement.  Each call advances a process-wide stream counter so repeated calls differ,
like repeated calls of the reference do.
"""
import numpy as np
import torch

from .. import ops

_state = {"seed": 1, "stream": 0}


def seed(value):
    """The reference never calls srand(); this is an extension for reproducible runs."""
    _state["seed"] = int(value)
    _state["stream"] = 0


def _validate(high, size, replace, p, exclusion):
    if size <= 0:
        raise ValueError("'size' must be a positive integer.")                       # pyx:23-24
    if not isinstance(replace, bool):
        raise TypeError("'replace' must be bool.")                                   # pyx:26-27
    if p is not None:
        raise NotImplementedError                                                    # pyx:29-30
    n_excl = len(exclusion) if exclusion is not None else 0
    if exclusion is not None and high <= n_excl:
        raise ValueError("The number of 'exclusion' is greater than 'high'.")        # pyx:32-33
    if replace is False and (high - n_excl <= size):
        raise ValueError("There is not enough integers to be sampled.")              # pyx:36-37


def batch_randint_choice(high, size, replace=True, p=None, exclusion=None):
    """-> list of per-row results: int when size[r] == 1 else list[int] (pyx:59-62, 64-89)."""
    if p is not None:
        raise NotImplementedError
    if exclusion is not None and len(size) != len(exclusion):
        raise ValueError("The shape of 'exclusion' is not compatible with the shape of 'size'!")
    sizes = [int(s) for s in size]
    for r, s in enumerate(sizes):
        _validate(high, s, replace, None, exclusion[r] if exclusion is not None else None)
    optr = np.zeros(len(sizes) + 1, dtype=np.int64)
    optr[1:] = np.cumsum(sizes)
    d_optr = torch.from_numpy(optr).cuda()
    eptr = eidx = None
    if exclusion is not None:
        rows = [np.unique(np.asarray(list(e), dtype=np.int32)) for e in exclusion]
        ep = np.zeros(len(rows) + 1, dtype=np.int64)
        ep[1:] = np.cumsum([len(r) for r in rows])
        ei = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
        eptr, eidx = torch.from_numpy(ep).cuda(), torch.from_numpy(np.ascontiguousarray(ei)).cuda()
        if eidx.numel() == 0:
            eidx = torch.zeros(1, dtype=torch.int32, device="cuda")
    _state["stream"] += 1
    flat = ops.batch_randint_choice(high, d_optr, int(optr[-1]), replace, eptr, eidx, _state["seed"],
                                    _state["stream"]).cpu().numpy()
    out = []
    for r, s in enumerate(sizes):
        row = flat[optr[r]:optr[r + 1]]
        out.append(int(row[0]) if s == 1 else row.tolist())
    return out


def randint_choice(high, size=1, replace=True, p=None, exclusion=None):
    _validate(high, size, replace, p, exclusion)
    return batch_randint_choice(high, [size], replace=replace, p=None,
                                exclusion=None if exclusion is None else [exclusion])[0]
