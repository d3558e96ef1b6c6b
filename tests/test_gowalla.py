"""BASELINE configs[2] on its REAL data: dataset/gowalla.{train,test} as loaded by the reference
(tests/golden/gowalla_split.npz) -- the oracle's and the product's adjacency against the
reference's own create_adj_mat('pre') (LightGCN.py:35-78), the oracle evaluator against the
reference's ProxyEvaluator strings."""
import zlib

import numpy as np

import oracle
from conftest import gowalla_tables, parse_result_string
from oracle import tf_math


def test_split_is_the_reference_split(gowalla):
    g, k = gowalla, gowalla["kat"]
    assert (g["num_users"], g["num_items"]) == (29858, 40981) == (k["num_users"], k["num_items"])
    assert len(g["train_indices"]) == 810128 == k["train_nnz"] and len(g["test_indices"]) == 217242 == k["test_nnz"]
    assert (np.diff(g["test_indptr"]) > 0).all()                      # all 29 858 users have test items
    for ptr, idx in ((g["train_indptr"], g["train_indices"]), (g["test_indptr"], g["test_indices"])):
        rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
        assert (np.diff(idx.astype(np.int64) + rows * 100000) > 0).all()   # ascending, duplicate-free rows


def _check_adj(A, g):
    k, a = g["kat"], g["adj"]
    A = A.tocsr(); A.sort_indices()
    assert list(A.shape) == k["adj_shape"] and A.nnz == k["adj_nnz"] == 1620256
    assert zlib.crc32(A.indices.astype(np.int32).tobytes()) == k["adj_indices_crc32"]
    data = A.data.astype(np.float32)
    assert zlib.crc32(data.tobytes()) == k["adj_data_crc32"]        # bit-identical values
    assert np.array_equal(np.diff(A.indptr).astype(np.int32), a["row_nnz"])
    assert np.array_equal(data[:256], a["data_head"]) and np.array_equal(data[-256:], a["data_tail"])
    assert np.allclose(np.asarray(A.sum(1)).ravel(), a["rowsum"], rtol=1e-5)


def test_oracle_adjacency_equals_the_reference_on_gowalla(gowalla):
    g = gowalla
    _check_adj(tf_math.lightgcn_adj(g["train_indptr"], g["train_indices"], g["num_users"], g["num_items"], "pre"), g)


def test_product_adjacency_equals_the_reference_on_gowalla(gowalla):
    from neurec_b200.model.general_recommender.LightGCN import bipartite_adjacency
    g = gowalla
    users = np.repeat(np.arange(g["num_users"], dtype=np.int32), np.diff(g["train_indptr"]))
    A = bipartite_adjacency(users, g["train_indices"], g["num_users"], g["num_items"], "pre", verbose=False)
    _check_adj(A.astype(np.float32), g)


def test_oracle_evaluator_reproduces_the_reference_on_a_gowalla_slice(gowalla):
    """ProxyEvaluator (cpp backend, np.matmul predict) on 512 gowalla users, 40 981 items: the C
    restatement fed OpenBLAS scores prints the same string; fed its own FMA-chain scores it stays
    within 1e-5 (north_star's NDCG bar)."""
    g = gowalla
    U, V = gowalla_tables(g)
    users = np.asarray(g["kat"]["subset_users"], dtype=np.int32)
    want = parse_result_string(g["kat"]["eval_subset_512"])
    tp = np.zeros(len(users) + 1, np.int64); tp[1:] = np.cumsum(np.diff(g["test_indptr"])[users])
    ti = np.concatenate([g["test_indices"][g["test_indptr"][u]:g["test_indptr"][u + 1]] for u in users])
    metric = [1, 2, 4, 3, 5]                                           # NeuRec.properties: Precision Recall NDCG MAP MRR
    rows_blas = []
    for off in range(0, len(users), 128):                              # test_batch_size=128
        ub = users[off:off + 128]
        s = np.matmul(U[ub], V.T).astype(np.float32)
        oracle.mask_train(s, ub, g["train_indptr"], g["train_indices"])
        rows_blas.append(oracle.evaluate_matrix(s, tp[off:off + len(ub) + 1] - tp[off], ti[tp[off]:tp[off + len(ub)]],
                                                metric, 20, thread_num=4))
    rows_blas = np.concatenate(rows_blas)
    got = np.mean(rows_blas, axis=0).reshape(5, 20)[:, [9, 19]].reshape(-1)
    text = "\t".join([("%.8f" % x).ljust(12) for x in got])
    assert text == g["kat"]["eval_subset_512"]
    rows = oracle.eval_mf(U, V, users, g["train_indptr"], g["train_indices"], tp, ti, metric, 20, thread_num=4)
    got2 = rows.astype(np.float64).mean(0).reshape(5, 20)[:, [9, 19]].reshape(-1)
    assert np.abs(got2 - want).max() < 1e-5
