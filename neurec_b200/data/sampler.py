"""PointwiseSampler / PairwiseSampler with the reference's interface (data/sampler.py:93-213)
on the device epoch (csrc/epoch.cu).

Python-visible behaviour is the reference's: construct once, iterate once per epoch (new
negatives and a new order every ``__iter__``), batches are python lists of length <= batch_size,
``len()`` is the number of batches, ``ValueError`` for ``neg_num <= 0``.  Differences, by design:
  * negatives come from the counter-based Philox kernel instead of the serial glibc ``rand()``
    loop -- same distribution (uniform over items the user has not interacted with, with
    replacement, aligned with the flattened positives), different stream;
  * the epoch order is a keyed bijection evaluated on the device (``nrc_shuffle_perm`` /
    ``nrc_epoch_build``) instead of ``np.random.permutation`` + per-sample python gathers
    (util/data_iterator.py:59,147-152 -- ~70 % of a reference sampler epoch);
  * ``device_epoch()`` hands the whole shuffled epoch to the training kernels as int32 CUDA
    tensors; MF trains straight from ``epoch_args()`` (sampling + shuffling + every step in one
    persistent launch, ``nrc_mf_epoch_fused``).
Like the reference's global ``rand()`` / ``np.random`` state, the stream position is process-wide:
every epoch drawn by ANY sampler takes the next epoch number, so a model that builds a new sampler
per epoch (MLP.py:100) still sees fresh negatives.
"""
import itertools

import numpy as np
import torch

from .. import ops

_EPOCH_COUNTER = itertools.count()      # process-wide stream position (the reference's global RNG state)


def _EPOCH_COUNTER_NEXT():
    """Next position of the process-wide stream (models that sample outside a Sampler object: SBPR)."""
    return next(_EPOCH_COUNTER)


def reseed(first_epoch=0):
    """Restart the process-wide epoch numbering (tests; the analogue of re-seeding np.random)."""
    global _EPOCH_COUNTER
    _EPOCH_COUNTER = itertools.count(int(first_epoch))


class Sampler(object):
    """Base class (data/sampler.py:9-21)."""

    def __len__(self):
        raise NotImplementedError

    def __iter__(self):
        raise NotImplementedError


def _generate_positive_items(user_pos_dict):
    # data/sampler.py:24-39
    if not isinstance(user_pos_dict, dict):
        raise TypeError("'user_pos_dict' must be a dict.")
    if not user_pos_dict:
        raise ValueError("'user_pos_dict' cannot be empty.")
    users = np.fromiter(user_pos_dict.keys(), dtype=np.int64)
    lens = np.fromiter((len(v) for v in user_pos_dict.values()), dtype=np.int64, count=len(users))
    users_list = np.repeat(users, lens).astype(np.int32)
    pos_items = np.concatenate([np.asarray(v, dtype=np.int32) for v in user_pos_dict.values()])
    return np.stack([users, lens], axis=1), users_list, pos_items


class _NegativeSamplerBase(Sampler):
    def __init__(self, dataset, neg_num, batch_size, shuffle, drop_last, seed):
        super().__init__()
        if neg_num <= 0:
            raise ValueError("'neg_num' must be a positive integer.")       # sampler.py:117-118,185-186
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.neg_num = neg_num
        self.item_num = dataset.num_items
        self.user_pos_dict = dataset.get_user_train_dict()
        self.user_pos_len, self._users_np, self._pos_np = _generate_positive_items(self.user_pos_dict)
        if int(self.user_pos_len[:, 1].max()) >= self.item_num:
            raise ValueError("The number of 'exclusion' is greater than 'high'.")  # pyx:32-33
        self.seed, self.epoch = int(seed), -1
        self._dev = None

    # device-resident train CSR (rows = sorted item lists) and flattened positives
    def _device_state(self):
        if self._dev is None:
            n_users = int(max(self.user_pos_dict.keys())) + 1
            ptr = np.zeros(n_users + 1, dtype=np.int64)
            for u, items in self.user_pos_dict.items():
                ptr[u + 1] = len(items)
            ptr = np.cumsum(ptr)
            idx = np.empty(int(ptr[-1]), dtype=np.int32)
            for u, items in self.user_pos_dict.items():
                idx[ptr[u]:ptr[u + 1]] = np.sort(np.asarray(items, dtype=np.int32))
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            self._dev = {"ptr": t(ptr), "idx": t(idx), "users": t(self._users_np), "pos": t(self._pos_np)}
        return self._dev

    def _next_epoch(self):
        self.epoch = next(_EPOCH_COUNTER)
        return self.epoch

    def _n_used(self):
        n = self._n_samples()
        return (n // self.batch_size) * self.batch_size if self.drop_last else n

    def epoch_args(self):
        """(train_indptr, train_indices, pos_users, pos_items) device tensors + the scalars
        nrc_mf_epoch_fused / nrc_epoch_build take; draws the next epoch number."""
        d = self._device_state()
        return d, dict(neg_num=self.neg_num, num_items=self.item_num, shuffle=bool(self.shuffle),
                       drop_last=bool(self.drop_last), seed=self.seed, epoch=self._next_epoch())

    def _device_epoch(self, pairwise):
        d, a = self.epoch_args()
        return ops.epoch_build(d["ptr"], d["idx"], d["users"], d["pos"], a["neg_num"], a["num_items"], pairwise,
                               a["shuffle"], a["seed"], a["epoch"], 0, self._n_used())

    def _n_samples(self):
        raise NotImplementedError

    def __len__(self):
        n = self._n_samples()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


class PairwiseSampler(_NegativeSamplerBase):
    """(users, pos_items, neg_items) batches; neg rows are lists when neg_num > 1."""

    def __init__(self, dataset, neg_num=1, batch_size=1024, shuffle=True, drop_last=False, seed=2018):
        super().__init__(dataset, neg_num, batch_size, shuffle, drop_last, seed)
        self.users_list, self.pos_items_list = self._users_np, self._pos_np

    def _n_samples(self):
        return len(self._users_np)

    def device_epoch(self):
        """Shuffled epoch as CUDA tensors: users [n], pos [n], neg [n] (or [n, neg_num])."""
        users, pos, neg = self._device_epoch(True)
        return users, pos, (neg.view(-1) if self.neg_num == 1 else neg)

    def __iter__(self):
        users, pos, neg = (t.cpu().numpy() for t in self.device_epoch())
        for off in range(0, len(users), self.batch_size):
            sl = slice(off, off + self.batch_size)
            yield users[sl].tolist(), pos[sl].tolist(), neg[sl].tolist()


class PointwiseSampler(_NegativeSamplerBase):
    """(users, items, labels) batches: positives labelled 1.0, negatives 0.0; the k-th negatives
    of all positives are contiguous before shuffling (data/sampler.py:121-147)."""

    def __init__(self, dataset, neg_num=1, batch_size=1024, shuffle=True, drop_last=False, seed=2018):
        super().__init__(dataset, neg_num, batch_size, shuffle, drop_last, seed)
        self.pos_items_list = self._pos_np
        self.users_list = np.tile(self._users_np, self.neg_num + 1)
        n_pos = len(self._pos_np)
        self.all_labels = np.concatenate([np.ones(n_pos, np.float32), np.zeros(n_pos * self.neg_num, np.float32)])

    def _n_samples(self):
        return len(self._users_np) * (self.neg_num + 1)

    def device_epoch(self):
        """Shuffled epoch as CUDA tensors: users [n] i32, items [n] i32, labels [n] f32."""
        return self._device_epoch(False)

    def __iter__(self):
        users, items, labels = (t.cpu().numpy() for t in self.device_epoch())
        for off in range(0, len(users), self.batch_size):
            sl = slice(off, off + self.batch_size)
            yield users[sl].tolist(), items[sl].tolist(), labels[sl].tolist()
