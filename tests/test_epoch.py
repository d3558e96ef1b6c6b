"""The device epoch order (csrc/epoch.cuh) through its CPU restatement: the keyed bijection that
replaces RandomSampler's np.random.permutation (util/data_iterator.py:45-63) must be a permutation
for every n, differ between epochs and seeds, and look like a random permutation statistically.
The GPU kernels are pinned bit for bit on this restatement in tests/test_gpu_epoch.py."""
import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 256, 257, 1000, 4097, 65536, 80367, 401835])
def test_is_a_permutation_for_every_n(n):
    p = oracle.shuffle_perm(n, 2018, 3)
    assert p.dtype == np.int64 and len(p) == n
    assert np.array_equal(np.sort(p), np.arange(n))


def test_identity_without_shuffle_and_new_order_every_epoch():
    n = 5000
    assert np.array_equal(oracle.shuffle_perm(n, 1, 0, shuffle=False), np.arange(n))
    a, b, c = oracle.shuffle_perm(n, 1, 0), oracle.shuffle_perm(n, 1, 1), oracle.shuffle_perm(n, 2, 0)
    assert np.array_equal(a, oracle.shuffle_perm(n, 1, 0))              # a pure function of (n, seed, epoch)
    for x, y in ((a, b), (a, c), (b, c)):
        assert (x == y).mean() < 0.01                                    # ~1/n fixed points in common
    assert (a == np.arange(n)).mean() < 0.01


def test_position_statistics_look_like_a_random_permutation():
    n, trials = 157, 4000
    first = np.array([oracle.shuffle_perm(n, 7, e)[0] for e in range(trials)])
    counts = np.bincount(first, minlength=n)
    chi2 = ((counts - trials / n) ** 2 / (trials / n)).sum()
    assert chi2 < n + 6 * np.sqrt(2 * n)                                 # chi-square(n-1): mean n-1, sd sqrt(2n)
    # where element 0 ends up is uniform too
    where = np.array([int(np.nonzero(oracle.shuffle_perm(n, 9, e) == 0)[0][0]) for e in range(trials)])
    counts = np.bincount(where, minlength=n)
    chi2 = ((counts - trials / n) ** 2 / (trials / n)).sum()
    assert chi2 < n + 6 * np.sqrt(2 * n)


def test_no_local_structure_survives():
    n = 80367
    p = oracle.shuffle_perm(n, 2018, 0)
    d = np.diff(p)
    assert (d == 1).mean() < 5.0 / n + 1e-4                              # consecutive inputs do not stay adjacent
    assert abs(np.corrcoef(np.arange(n), p)[0, 1]) < 0.02
    # a batch of 512 consecutive positions covers the id range evenly (what shuffling is for)
    batch = p[:512]
    assert np.abs(batch.mean() / n - 0.5) < 0.06


def test_epoch_build_layouts(ml100k):
    tp, ti = ml100k["train_indptr"], ml100k["train_indices"]
    nu, ni = ml100k["num_users"], ml100k["num_items"]
    pos_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(tp))
    u, i, j = oracle.epoch_build(tp, ti, pos_users, ti, 1, ni, True, True, 2018, 0)
    assert len(u) == len(ti) and j.shape == (len(ti), 1)
    # every (user, positive) pair exactly once; negatives never in the user's train row
    key = u.astype(np.int64) * ni + i
    assert np.array_equal(np.sort(key), np.sort(pos_users.astype(np.int64) * ni + ti))
    rows = {int(x): set(ti[tp[x]:tp[x + 1]].tolist()) for x in range(nu)}
    assert all(int(b) not in rows[int(a)] for a, b in zip(u[:5000], j[:5000, 0]))
    # pointwise: positives labelled 1, neg_num negatives per positive labelled 0
    u2, i2, l2 = oracle.epoch_build(tp, ti, pos_users, ti, 4, ni, False, True, 2018, 1)
    assert len(u2) == 5 * len(ti) and l2.dtype == np.float32 and l2.sum() == len(ti)
    pk = u2[l2 == 1].astype(np.int64) * ni + i2[l2 == 1]
    assert np.array_equal(np.sort(pk), np.sort(pos_users.astype(np.int64) * ni + ti))
    assert np.array_equal(np.bincount(u2[l2 == 0], minlength=nu), 4 * np.diff(tp))
    # unshuffled pointwise layout is the reference's: positives, then k-major negatives (sampler.py:139-141)
    u3, i3, l3 = oracle.epoch_build(tp, ti, pos_users, ti, 2, ni, False, False, 5, 0)
    neg = oracle.philox_sample_negatives(tp, ti, pos_users, 2, ni, 5, 0)
    n = len(ti)
    assert np.array_equal(i3[:n], ti) and np.array_equal(i3[n:2 * n], neg[:, 0]) and np.array_equal(i3[2 * n:], neg[:, 1])
    assert np.array_equal(u3, np.tile(pos_users, 3))


def test_unshuffled_epoch_layout_equals_the_reference_samplers(ml100k):
    """One unshuffled epoch of the REAL PointwiseSampler (neg_num=2) and PairwiseSampler (neg_num=3) on the ml-100k split
    (tests/golden/kat_sampler_layout.json): the restated epoch (oracle.epoch_build, which the device kernels match bit for
    bit) puts the same users, labels and positives in the same places; its negatives (another stream) obey the same rule."""
    import json
    import os
    import zlib
    import oracle
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_sampler_layout.json")) as f:
        kat = json.load(f)
    d = ml100k
    tp, ti, ni = d["train_indptr"], d["train_indices"], d["num_items"]
    pos_users = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(tp))
    crc = lambda a, t: zlib.crc32(np.asarray(a, dtype=t).tobytes())
    u, i, lab = oracle.epoch_build(tp, ti, pos_users, ti, 2, ni, False, False, 2018, 0)
    n_pos = len(pos_users)
    assert len(u) == kat["n"] and n_pos == kat["n_pos"]
    assert crc(u, np.int32) == kat["users_crc32"] and crc(lab, np.float32) == kat["labels_crc32"]
    assert crc(i[:n_pos], np.int32) == kat["pos_items_crc32"]
    rows = [set(ti[tp[x]:tp[x + 1]].tolist()) for x in range(d["num_users"])]
    assert all(int(i[e]) not in rows[u[e]] for e in range(n_pos, len(u)))
    pu, pp, pn = oracle.epoch_build(tp, ti, pos_users, ti, 3, ni, True, False, 2018, 0)
    assert len(pu) == kat["pair_n"] and list(pn.shape) == kat["pair_neg_shape"]
    assert crc(pu, np.int32) == kat["pair_users_crc32"] and crc(pp, np.int32) == kat["pair_pos_crc32"]
    assert all(int(j) not in rows[x] for x, r in zip(pu, pn) for j in r)
