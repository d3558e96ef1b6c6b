"""TF-free helpers of the reference's util/tool.py that the hot path's callers use."""
import time
from functools import wraps
from inspect import signature

import numpy as np


def csr_to_user_dict(train_matrix):
    """{row: ascending item list} for non-empty rows (reference util/tool.py:56-65)."""
    m = train_matrix.tocsr()
    m.sort_indices()
    ptr, idx = m.indptr, m.indices
    return {u: idx[ptr[u]:ptr[u + 1]].tolist() for u in range(m.shape[0]) if ptr[u + 1] > ptr[u]}


def csr_to_user_dict_bytime(time_matrix, train_matrix):
    """Items of each user ordered by interaction time (reference util/tool.py:68-76)."""
    out = {}
    for u, items in csr_to_user_dict(train_matrix).items():
        out[u] = np.array(sorted(items, key=lambda x: time_matrix[u, x]), dtype=np.int32).tolist()
    return out


def typeassert(*type_args, **type_kwargs):
    """Argument type check decorator -> TypeError (reference util/tool.py:136-150)."""
    def decorate(func):
        sig = signature(func)
        bound_types = sig.bind_partial(*type_args, **type_kwargs).arguments

        @wraps(func)
        def wrapper(*args, **kwargs):
            for name, value in sig.bind(*args, **kwargs).arguments.items():
                if name in bound_types and not isinstance(value, bound_types[name]):
                    raise TypeError("Argument {} must be {}".format(name, bound_types[name]))
            return func(*args, **kwargs)
        return wrapper
    return decorate


def timer(func):
    """Prints the wall time of `func` (reference util/tool.py:203-213)."""
    @wraps(func)
    def wrapper(*args, **kwargs):
        t0 = time.time()
        result = func(*args, **kwargs)
        print("%s function cost: %fs" % (func.__name__, time.time() - t0))
        return result
    return wrapper


def pad_sequences(sequences, value=0., max_len=None, padding="post", truncating="post", dtype=np.int32):
    """Reference util/tool.py:158-200."""
    if max_len is None:
        max_len = int(np.max([len(x) for x in sequences]))
    x = np.full([len(sequences), max_len], value, dtype=dtype)
    for i, s in enumerate(sequences):
        if not len(s):
            continue
        if truncating == "pre":
            trunc = s[-max_len:]
        elif truncating == "post":
            trunc = s[:max_len]
        else:
            raise ValueError('Truncating type "%s" not understood' % truncating)
        if padding == "post":
            x[i, :len(trunc)] = trunc
        elif padding == "pre":
            x[i, -len(trunc):] = trunc
        else:
            raise ValueError('Padding type "%s" not understood' % padding)
    return x


def randint_choice(high, size=None, replace=True, p=None, exclusion=None):
    """numpy sampler used by Dataset for test negatives (reference util/tool.py:120-133)."""
    a = np.arange(high)
    if exclusion is not None:
        p = np.ones_like(a) if p is None else np.array(p, copy=True)
        p = p.flatten().astype(np.float64)
        p[exclusion] = 0
        p = p / np.sum(p)
    return np.random.choice(a, size=size, replace=replace, p=p)
