"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the neurec_b200 hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``neurec_b200``) never
does; it raises when its CUDA library is missing instead of falling back to anything here.

Three things live here:

* ``neurec_oracle.c``  -- plain-C restatement of the reference's compiled path
  (evaluator, metrics, libstdc++ partial_sort_copy tie order, libc-rand sampler) and of the
  product's counter-based streams (Philox negatives, keyed-bijection order, SBPR draws, split keys).
* ``tf_math.py``       -- numpy fp32 restatement of the TF-1.12 graphs (MF/BPR, pointwise,
  MLP/NeuMF, LightGCN, NGCF, APR, SBPR, SpectralCF) and of the TF optimizers; parity UNPINNED at
  the TensorFlow boundary (no TF offline), gradients pinned by finite differences.
* numpy restatements in this file of the data side: SBPR's social-item sets (crc-identical to the
  real reference's on Ciao), interactions -> CSR, the per-user train/test split (bit-identical to the
  real reference's split_by_ratio / split_by_loo on ml-100k).
* ``_ref/``            -- the REAL reference: its C++ headers and .pyx files compiled from
  where they lie under /root/reference by ``oracle/Makefile`` (``make ref``).  Used to pin
  the restatements and, in bench.py, as the ``"kind": "reference"`` CPU baseline.

Parity status: pinned (see the header of neurec_oracle.c and tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import importlib.abc
import importlib.util
import os
import subprocess
import sys
import sysconfig
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("NRC_REFERENCE_ROOT", "/root/reference")
_EXT = sysconfig.get_config_var("EXT_SUFFIX")

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_i64p = ctypes.POINTER(ctypes.c_int64)


def build(ref: bool = True) -> None:
    """Compile the C restatement and (when /root/reference exists) oracle/_ref."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE, f"REF={REF_ROOT}"] + targets, check=True)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_c_f32p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_c_i32p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_c_i64p)


_LIB = None


def lib() -> ctypes.CDLL:
    """ctypes handle of the C restatement (builds it on first use when missing)."""
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "libneurec_oracle.so")
        src = os.path.join(HERE, "neurec_oracle.c")
        if not os.path.isfile(path) or os.path.getmtime(src) > os.path.getmtime(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        L.orc_llrand.restype = ctypes.c_ulonglong
        _LIB = L
    return _LIB


_REF = False


def ref_lib():
    """ctypes handle of the compiled reference headers, or None when not built."""
    global _REF
    if _REF is False:
        path = os.path.join(HERE, "_ref", "libneurec_ref.so")
        _REF = ctypes.CDLL(path) if os.path.isfile(path) else None
    return _REF


# ----------------------------------------------------------------------------------------
# numpy-facing wrappers
# ----------------------------------------------------------------------------------------
METRIC_IDS = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}


def dict_to_csr(d, num_rows):
    """{row: [sorted ints]} -> (indptr int64[num_rows+1], indices int32)."""
    indptr = np.zeros(num_rows + 1, dtype=np.int64)
    for u, items in d.items():
        indptr[u + 1] = len(items)
    indptr = np.cumsum(indptr)
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for u, items in d.items():
        indices[indptr[u]:indptr[u + 1]] = np.sort(np.asarray(list(items), dtype=np.int32))
    return indptr, indices


def lists_to_csr(rows):
    """[iterable of ints per row] -> CSR with sorted, duplicate-free rows."""
    rows = [np.unique(np.asarray(list(r), dtype=np.int32)) for r in rows]
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)
    return indptr, indices


def evaluate_matrix(scores, test_indptr, test_indices, metric, top_k, thread_num=1,
                    return_ranks=False, impl="oracle"):
    """eval_score_matrix (cpp_evaluator.pyx:28-42) on the C restatement (impl="oracle") or
    on the real reference headers (impl="reference")."""
    scores, sp = _f32(scores)
    B, N = scores.shape
    test_indptr, ip = _i64(test_indptr)
    test_indices, xp = _i32(test_indices)
    metric, mp = _i32(metric)
    out = np.zeros((B, len(metric) * top_k), dtype=np.float32)
    if impl == "reference":
        R = ref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref is not built")
        R.ref_cpp_evaluate_matrix(sp, ctypes.c_int(N), ctypes.c_int(B), ip, xp, mp,
                                  ctypes.c_int(len(metric)), ctypes.c_int(top_k),
                                  ctypes.c_int(thread_num), out.ctypes.data_as(_c_f32p))
        return out
    ranks = np.zeros((B, top_k), dtype=np.int32)
    rc = lib().orc_evaluate_matrix(sp, ctypes.c_int(N), ctypes.c_int(B), ip, xp, mp,
                                   ctypes.c_int(len(metric)), ctypes.c_int(top_k),
                                   ctypes.c_int(thread_num), out.ctypes.data_as(_c_f32p),
                                   ranks.ctypes.data_as(_c_i32p))
    if rc:
        raise ValueError("unknown metric id")
    return (out, ranks) if return_ranks else out


def arg_topk(scores, top_k, thread_num=1, impl="oracle"):
    """arg_topk (arg_topk.pyx:16-35)."""
    scores, sp = _f32(scores)
    U, N = scores.shape
    out = np.zeros((U, top_k), dtype=np.int32)
    if impl == "reference":
        ref_lib().ref_arg_top_k_2d(sp, ctypes.c_int(N), ctypes.c_int(U), ctypes.c_int(top_k),
                                   ctypes.c_int(thread_num), out.ctypes.data_as(_c_i32p))
    else:
        lib().orc_arg_topk_2d(sp, ctypes.c_int(N), ctypes.c_int(U), ctypes.c_int(top_k),
                              ctypes.c_int(thread_num), out.ctypes.data_as(_c_i32p))
    return out


def mf_scores(U, V, users, thread_num=1):
    """fp32 FMA-chain scores [len(users), num_items] (the oracle's definition of predict)."""
    U, up = _f32(U)
    V, vp = _f32(V)
    users, usp = _i32(users)
    out = np.empty((len(users), V.shape[0]), dtype=np.float32)
    lib().orc_mf_scores(up, vp, usp, ctypes.c_int(len(users)), ctypes.c_int(V.shape[0]),
                        ctypes.c_int(V.shape[1]), ctypes.c_int(thread_num),
                        out.ctypes.data_as(_c_f32p))
    return out


def mask_train(scores, users, train_indptr, train_indices):
    """In place: scores[b, train(users[b])] = -inf (uni_evaluator.py:140-143)."""
    assert scores.dtype == np.float32 and scores.flags.c_contiguous
    users, usp = _i32(users)
    train_indptr, ip = _i64(train_indptr)
    train_indices, xp = _i32(train_indices)
    lib().orc_mask_train(scores.ctypes.data_as(_c_f32p), usp, ctypes.c_int(len(users)),
                         ctypes.c_int(scores.shape[1]), ip, xp)
    return scores


def eval_mf(U, V, users, train_indptr, train_indices, test_indptr_b, test_indices_b, metric,
            top_k, thread_num=1, return_ranks=False):
    """predict -> mask -> eval for one batch of users; test CSR is per batch row."""
    s = mf_scores(U, V, users, thread_num)
    mask_train(s, users, train_indptr, train_indices)
    return evaluate_matrix(s, test_indptr_b, test_indices_b, metric, top_k, thread_num,
                           return_ranks=return_ranks)


def batch_randint_choice(high, sizes, replace=True, exclusion_csr=None):
    """batch_randint_choice (random_choice.pyx:64-89) on glibc rand(); flat int32 result."""
    sizes, sp = _i32(sizes)
    out = np.empty(int(sizes.sum()), dtype=np.int32)
    if exclusion_csr is not None:
        ei, eip = _i64(exclusion_csr[0])
        ex, exp_ = _i32(exclusion_csr[1])
    else:
        eip, exp_ = None, None
    rc = lib().orc_batch_randint_choice(ctypes.c_int(high), sp, ctypes.c_int(len(sizes)),
                                        ctypes.c_int(1 if replace else 0), eip, exp_,
                                        out.ctypes.data_as(_c_i32p))
    if rc:
        raise ValueError({-1: "'size' must be a positive integer.",
                          -2: "The number of 'exclusion' is greater than 'high'.",
                          -3: "There is not enough integers to be sampled."}[rc])
    return out


def philox_sample_negatives(train_indptr, train_indices, users, neg_num, num_items, seed,
                            stream_id, first_index=0):
    """CPU restatement of the product's Philox sampler (bit-exact target for the kernel)."""
    tp, tpp = _i64(train_indptr)
    ti, tip = _i32(train_indices)
    us, usp = _i32(users)
    out = np.empty((len(us), neg_num), dtype=np.int32)
    lib().orc_philox_sample_negatives(tpp, tip, usp, ctypes.c_int64(len(us)), ctypes.c_int(neg_num),
                                      ctypes.c_int(num_items), ctypes.c_uint64(seed),
                                      ctypes.c_uint64(stream_id), ctypes.c_int64(first_index),
                                      out.ctypes.data_as(_c_i32p))
    return out


def shuffle_perm(n, seed, epoch, shuffle=True):
    """CPU restatement of the product's epoch order (csrc/epoch.cuh): int64 [n] permutation of
    range(n) -- the stand-in for RandomSampler's np.random.permutation (data_iterator.py:45-63)."""
    out = np.empty(int(n), dtype=np.int64)
    lib().orc_feistel_perm(ctypes.c_int64(int(n)), ctypes.c_int(1 if shuffle else 0), ctypes.c_uint64(seed),
                           ctypes.c_uint64(epoch), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return out


def epoch_build(train_indptr, train_indices, pos_users, pos_items, neg_num, num_items, pairwise, shuffle,
                seed, epoch):
    """One epoch of Pairwise / PointwiseSampler (data/sampler.py:189-206 / 121-147) in the product's
    order and with the product's negatives: (users, items, third); third = int32 [n, neg_num]
    negatives (pairwise) or float32 [n] labels (pointwise, positives then k-major negatives)."""
    pos_users = np.ascontiguousarray(pos_users, dtype=np.int32)
    pos_items = np.ascontiguousarray(pos_items, dtype=np.int32)
    n_pos = len(pos_users)
    neg = philox_sample_negatives(train_indptr, train_indices, pos_users, neg_num, num_items, seed, epoch)
    if pairwise:
        perm = shuffle_perm(n_pos, seed, epoch, shuffle)
        return pos_users[perm], pos_items[perm], neg[perm]
    users = np.tile(pos_users, neg_num + 1)
    items = np.concatenate([pos_items, neg.T.reshape(-1)])
    labels = np.concatenate([np.ones(n_pos, np.float32), np.zeros(n_pos * neg_num, np.float32)])
    perm = shuffle_perm(len(users), seed, epoch, shuffle)
    return users[perm], items[perm], labels[perm]


def philox_batch_choice(high, out_indptr, replace=True, exclusion_csr=None, seed=0, stream_id=0):
    op, opp = _i64(out_indptr)
    out = np.empty(int(op[-1]), dtype=np.int32)
    if exclusion_csr is not None:
        ei, eip = _i64(exclusion_csr[0])
        ex, exp_ = _i32(exclusion_csr[1])
    else:
        eip, exp_ = None, None
    lib().orc_philox_batch_choice(ctypes.c_int(high), opp, ctypes.c_int(len(op) - 1),
                                  ctypes.c_int(1 if replace else 0), eip, exp_,
                                  ctypes.c_uint64(seed), ctypes.c_uint64(stream_id),
                                  out.ctypes.data_as(_c_i32p))
    return out


def social_items_csr(train_indptr, train_indices, trust_indptr, trust_indices):
    """SBPR._get_SocialItemsSet (social_recommender/SBPR.py:39-49) as a CSR with ascending rows: for every user
    the items its trusted users interacted with, minus its own."""
    tp = np.asarray(train_indptr, np.int64); ti = np.asarray(train_indices, np.int32)
    fp = np.asarray(trust_indptr, np.int64); fi = np.asarray(trust_indices, np.int32)
    rows = []
    for u in range(len(tp) - 1):
        own = set(ti[tp[u]:tp[u + 1]].tolist())
        if not own:                        # the reference iterates train_dict: users without train items are skipped
            rows.append(np.zeros(0, np.int32)); continue
        items = {int(i) for f in fi[fp[u]:fp[u + 1]] for i in ti[tp[f]:tp[f + 1]] if int(i) not in own}
        rows.append(np.asarray(sorted(items), np.int32))
    return lists_to_csr(rows)


def sbpr_epoch_build(train_indptr, train_indices, social_indptr, social_indices, trust_indptr, trust_indices,
                     pos_users, pos_items, num_items, shuffle, seed, epoch):
    """One epoch of SBPR._get_pairwise_all_data + DataIterator (SBPR.py:103-149) in the product's order and with
    the product's draws: (users, pos, social, neg, suk)."""
    tp, tpp = _i64(train_indptr); ti, tip = _i32(train_indices)
    sp, spp = _i64(social_indptr); si, sip = _i32(social_indices)
    fp, fpp = _i64(trust_indptr); fi, fip = _i32(trust_indices)
    us, usp = _i32(pos_users)
    pos_items = np.ascontiguousarray(pos_items, dtype=np.int32)
    n = len(us)
    soc = np.empty(n, np.int32); neg = np.empty(n, np.int32); suk = np.empty(n, np.float32)
    lib().orc_sbpr_sample(tpp, tip, spp, sip, fpp, fip, usp, ctypes.c_int64(n), ctypes.c_int(int(num_items)),
                          ctypes.c_uint64(seed), ctypes.c_uint64(epoch), soc.ctypes.data_as(_c_i32p),
                          neg.ctypes.data_as(_c_i32p), suk.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    perm = shuffle_perm(n, seed, epoch, shuffle)
    return us[perm], pos_items[perm], soc[perm], neg[perm], suk[perm]


def csr_from_coo(rows, cols, num_rows):
    """Interactions -> CSR with ascending duplicate-free rows (dataset.py:288-296 + tool.py:56-65)."""
    rows = np.asarray(rows, np.int64); cols = np.asarray(cols, np.int64)
    key = np.unique(rows * (int(cols.max()) + 1 if len(cols) else 1) + cols)
    width = int(cols.max()) + 1 if len(cols) else 1
    r, c = key // width, key % width
    indptr = np.zeros(num_rows + 1, np.int64)
    np.add.at(indptr, r + 1, 1)
    return np.cumsum(indptr), c.astype(np.int32)


def split_interactions(users, keys, num_users, mode="ratio", ratio=0.8, seed=0):
    """split_by_ratio / split_by_loo (data/utils.py:59-106) per interaction: 1 = train, 0 = test.  keys = interaction
    times (by_time=True) or None (by_time=False: the product's counter-based random order, seed-keyed); ties by input
    position; cut = ceil(ratio * n_u), or n_u - 1 when n_u > 3 for leave-one-out."""
    import math
    users = np.asarray(users, np.int64)
    n = len(users)
    if keys is None:
        k = np.empty(n, np.uint64)
        lib().orc_philox_words(ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.c_uint64(0x53504C4954000000),
                               k.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    else:
        k = (np.asarray(keys, np.int64).view(np.uint64) ^ np.uint64(1 << 63))
    order = np.lexsort((np.arange(n), k, users))            # by user, then key, then input position
    out = np.zeros(n, np.int32)
    counts = np.bincount(users, minlength=num_users)
    start = 0
    for u in range(num_users):
        c = int(counts[u])
        if c:
            cut = math.ceil(ratio * c) if mode == "ratio" else (c if c <= 3 else c - 1)
            out[order[start:start + cut]] = 1
        start += c
    return out


# ----------------------------------------------------------------------------------------
# Importing the REAL Python reference (build container only; never on the GPU box)
# ----------------------------------------------------------------------------------------
_REF_EXT = {
    "util.cython.random_choice": "random_choice",
    "util.cython.tools": "tools",
    "util.cython.arg_topk": "arg_topk",
    "evaluator.backend.cpp.cpp_evaluator": "cpp_evaluator",
}


class _RefExtFinder(importlib.abc.MetaPathFinder):
    """Resolves the reference's four Cython extension modules to oracle/_ref/*.so."""

    def find_spec(self, fullname, path, target=None):
        stem = _REF_EXT.get(fullname)
        if stem is None:
            return None
        so = os.path.join(HERE, "_ref", stem + _EXT)
        if not os.path.isfile(so):
            return None
        return importlib.util.spec_from_file_location(fullname, so)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "main.py")) and \
        os.path.isfile(os.path.join(HERE, "_ref", "cpp_evaluator" + _EXT))


def import_reference(scratch="/tmp/nrc_ref_cwd"):
    """Make ``import util / data / evaluator`` resolve to the unmodified reference.

    Two shims: a stub ``tensorflow`` module (TF 1.12 is not
    installable here; only TF-free pieces are used) and ``collections.Iterable``.
    A scratch working directory holds symlinks to NeuRec.properties, conf/ and the dataset
    files so the reference can write its split cache and logs without touching
    /root/reference.  Returns the scratch directory (the caller should chdir into it).
    """
    if not reference_available():
        raise RuntimeError("reference or oracle/_ref not available")
    import collections
    import collections.abc
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    if not any(isinstance(f, _RefExtFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefExtFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.makedirs(os.path.join(scratch, "dataset"), exist_ok=True)
    for name in ("NeuRec.properties", "conf"):
        dst = os.path.join(scratch, name)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF_ROOT, name), dst)
    for name in os.listdir(os.path.join(REF_ROOT, "dataset")):
        dst = os.path.join(scratch, "dataset", name)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF_ROOT, "dataset", name), dst)
    return scratch
