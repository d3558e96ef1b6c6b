"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU path, end to end, for timing and
for cross-checks.  Used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm and by
tests; never by the product.

The GPU box has no /root/reference, so this module composes
  * the REAL reference code that is compiled and therefore travels in ``oracle/_ref``
    (``libneurec_ref.so`` = evaluate.h + metric.h + arg_topk.h; ``random_choice*.so`` = the
    Cython sampler RNG), when present  -> ``kind = "reference"``
  * with plain-Python / numpy restatements of the reference's Python glue, written to do the
    same work in the same way (python lists, per-batch loops) so that the timing is comparable:
      pairwise_epoch / pointwise_epoch   data/sampler.py:24-39, 71-90, 121-147, 198-206 and
                                         util/data_iterator.py:45-63, 133-155 (np.random.permutation
                                         shuffle, list gather, zip(*) transposition)
      evaluate                           evaluator/backend/cpp/uni_evaluator.py:101-157
      train step                         oracle/tf_math.py (TensorFlow 1.12 is not installable).
"""
from __future__ import annotations

import importlib.util
import os
import sysconfig
import time

import numpy as np

import oracle
from oracle import tf_math

_EXT = sysconfig.get_config_var("EXT_SUFFIX")


def _ref_random_choice():
    """The reference's compiled util/cython/random_choice module, or None."""
    so = os.path.join(oracle.HERE, "_ref", "random_choice" + _EXT)
    if not os.path.isfile(so):
        return None
    spec = importlib.util.spec_from_file_location("random_choice", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_RC = False


def batch_randint_choice(high, size, replace=True, exclusion=None):
    global _RC
    if _RC is False:
        _RC = _ref_random_choice()
    if _RC is not None:
        return _RC.batch_randint_choice(high, size, replace=replace, exclusion=exclusion)
    csr = oracle.lists_to_csr(exclusion) if exclusion is not None else None
    flat = oracle.batch_randint_choice(high, size, replace, csr)
    out, off = [], 0
    for s in size:
        out.append(flat[off:off + s].tolist() if s > 1 else int(flat[off]))
        off += s
    return out


def sampler_kind():
    global _RC
    if _RC is False:
        _RC = _ref_random_choice()
    return "reference" if _RC is not None else "port"


def user_dict(indptr, indices):
    """util/tool.py:56-65 csr_to_user_dict."""
    return {u: indices[indptr[u]:indptr[u + 1]].tolist()
            for u in range(len(indptr) - 1) if indptr[u + 1] > indptr[u]}


def _iterate(arrays, batch_size, shuffle, drop_last=False):
    """util/data_iterator.py: RandomSampler (np.random.permutation) + BatchSampler +
    per-sample python gather + zip(*) transposition."""
    n = len(arrays[0])
    order = np.random.permutation(n).tolist() if shuffle else range(n)
    batch = []
    for idx in order:
        batch.append(idx)
        if len(batch) == batch_size:
            rows = [[a[i] for a in arrays] for i in batch]
            yield [list(s) for s in zip(*rows)]
            batch = []
    if batch and not drop_last:
        rows = [[a[i] for a in arrays] for i in batch]
        yield [list(s) for s in zip(*rows)]


class PairwiseSamplerPort:
    """data/sampler.py:158-213."""

    def __init__(self, train_dict, num_items, neg_num=1, batch_size=1024, shuffle=True):
        self.d, self.num_items, self.neg_num = train_dict, num_items, neg_num
        self.batch_size, self.shuffle = batch_size, shuffle
        self.user_pos_len, self.users_list, self.pos_items_list = [], [], []
        for user, pos in train_dict.items():                      # sampler.py:24-39
            self.user_pos_len.append([user, len(pos)])
            self.users_list.extend([user] * len(pos))
            self.pos_items_list.extend(pos)

    def _negatives(self):                                         # sampler.py:71-90
        users, n_pos = list(zip(*self.user_pos_len))
        neg = []
        for off in range(0, len(users), 1024):
            bu = users[off:off + 1024]
            bn = [n * self.neg_num for n in n_pos[off:off + 1024]]
            excl = [self.d[u] for u in bu]
            for items in batch_randint_choice(self.num_items, bn, replace=True, exclusion=excl):
                if isinstance(items, (list, tuple, np.ndarray)) or hasattr(items, "__iter__"):
                    items = list(items)
                    if self.neg_num > 1:
                        items = np.reshape(items, [-1, self.neg_num]).tolist()
                    neg.extend(items)
                else:
                    neg.append(items)
        return neg

    def __iter__(self):
        neg = self._negatives()
        yield from _iterate([self.users_list, self.pos_items_list, neg], self.batch_size, self.shuffle)

    def __len__(self):
        return (len(self.users_list) + self.batch_size - 1) // self.batch_size


class PointwiseSamplerPort(PairwiseSamplerPort):
    """data/sampler.py:93-155."""

    def __iter__(self):
        neg = np.array(self._negatives(), dtype=np.int32)
        neg = np.reshape(neg.T, [-1]).tolist()                    # sampler.py:139-141
        n_pos = len(self.pos_items_list)
        users = self.users_list * (self.neg_num + 1)
        items = self.pos_items_list + neg
        labels = [1.0] * n_pos + [0.0] * (n_pos * self.neg_num)
        yield from _iterate([users, items, labels], self.batch_size, self.shuffle)

    def __len__(self):
        n = len(self.users_list) * (self.neg_num + 1)
        return (n + self.batch_size - 1) // self.batch_size


def evaluate(U, V, train_dict, test_dict, metric_ids, top_k, batch_size=128, num_thread=8,
             predict=None):
    """UniEvaluator.evaluate, cpp/uni_evaluator.py:101-157 (rec.evaluate.neg == 0 branch), with
    MF.predict = np.matmul (MF.py:120-122).  Uses the real compiled evaluate.h when available.
    Returns (per-user result matrix, seconds spent in predict / mask / native)."""
    impl = "reference" if oracle.ref_lib() is not None else "oracle"
    users = list(test_dict.keys())
    out, t_pred, t_mask, t_nat = [], 0.0, 0.0, 0.0
    for off in range(0, len(users), batch_size):
        bu = users[off:off + batch_size]
        t0 = time.perf_counter()
        score = predict(bu) if predict is not None else np.matmul(U[bu], V.T)
        score = np.ascontiguousarray(score, dtype=np.float32)
        t1 = time.perf_counter()
        for idx, u in enumerate(bu):                               # uni_evaluator.py:140-143
            if u in train_dict and len(train_dict[u]) > 0:
                score[idx][train_dict[u]] = -np.inf
        t2 = time.perf_counter()
        ip, ix = oracle.lists_to_csr([test_dict[u] for u in bu])   # pyx:32 list -> set marshalling
        out.append(oracle.evaluate_matrix(score, ip, ix, metric_ids, top_k, num_thread, impl=impl))
        t3 = time.perf_counter()
        t_pred += t1 - t0; t_mask += t2 - t1; t_nat += t3 - t2
    return np.concatenate(out, axis=0), (t_pred, t_mask, t_nat), impl
