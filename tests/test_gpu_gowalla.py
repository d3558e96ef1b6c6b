"""BASELINE configs[2] on the GPU with its REAL data (dataset/gowalla 'given' split, fixture
tests/golden/gowalla_split.npz): LightGCN steps against the numpy/scipy restatement on the real
70 839-node / 1 620 256-nnz graph, and the full 29 858-user evaluation against the reference's
own ProxyEvaluator output and the C oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import gowalla_tables, parse_result_string
from oracle import tf_math

pytestmark = pytest.mark.gpu
METRIC_NAMES = ["Precision", "Recall", "NDCG", "MAP", "MRR"]          # NeuRec.properties order


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _graph(g):
    A = tf_math.lightgcn_adj(g["train_indptr"], g["train_indices"], g["num_users"], g["num_items"], "pre")
    csr = (dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data.astype(np.float32)))
    order = dev(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
    return A, csr, order


def test_lightgcn_steps_on_the_real_gowalla_graph(gowalla):
    """conf/LightGCN.properties (lr 0.01, reg 1e-3, d 64, bs 1024) + --n_layers=3: 4 steps on triplets
    of the product's own device epoch (sampler + shuffle), vs tf_math.LightGCNTrainer."""
    from neurec_b200 import ops
    g = gowalla
    nu, ni, dim, L, bs, steps = g["num_users"], g["num_items"], 64, 3, 1024, 4
    A, csr, order = _graph(g)
    assert A.shape == (70839, 70839) and A.nnz == 1620256
    rs = np.random.RandomState(4)
    lim = np.sqrt(6.0 / (nu + dim))
    e0 = rs.uniform(-lim, lim, (nu + ni, dim)).astype(np.float32)
    pos_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(g["train_indptr"]))
    u, i, j = ops.epoch_build(dev(g["train_indptr"]), dev(g["train_indices"]), dev(pos_users), dev(g["train_indices"]),
                              1, ni, True, True, 2018, 0, 0, bs * steps)
    wu, wi, wj = oracle.epoch_build(g["train_indptr"], g["train_indices"], pos_users, g["train_indices"], 1, ni, True,
                                    True, 2018, 0)
    assert np.array_equal(u.cpu().numpy(), wu[:bs * steps]) and np.array_equal(j.cpu().numpy(), wj[:bs * steps])
    tr = tf_math.LightGCNTrainer(A, e0, nu, L, 0.01, 1e-3)
    want = tr.epoch(wu[:bs * steps], wi[:bs * steps], wj[:bs * steps, 0], bs)
    de0 = dev(e0)
    z = lambda: torch.zeros_like(de0)
    m, v, ef, gf, ge, wa, wb = z(), z(), z(), z(), z(), z(), z()
    sl = torch.zeros(steps, 2, device="cuda")
    n = ops.lightgcn_train_epoch(csr, None, order, nu, ni, L, de0, m, v, u, i, j.view(-1), bs, 1e-3,
                                 tf_math.adam_lr_t(0.01, steps), [0.01, 0.9, 0.999, 1e-8], ef, gf, ge, (wa, wb), sl)
    assert n == steps
    assert np.allclose(sl.cpu().numpy(), want, rtol=2e-4)
    assert np.abs(de0.cpu().numpy() - tr.e0).max() < 5e-5
    # propagation of the trained table: within fp32 re-association of the exact scipy product
    prop, _ = tf_math.lightgcn_propagate(A, de0.cpu().numpy(), L)
    got = ops.lightgcn_propagate(csr[0], csr[1], csr[2], order, de0, L).cpu().numpy()
    assert np.abs(got - prop).max() < 1e-6


def _uni(g, top_k=(10, 20), batch_size=4096):
    from neurec_b200.evaluator.uni_evaluator import UniEvaluator
    tp, ti, sp, si = g["train_indptr"], g["train_indices"], g["test_indptr"], g["test_indices"]
    train = {u: ti[tp[u]:tp[u + 1]].tolist() for u in range(g["num_users"])}
    test = {u: si[sp[u]:sp[u + 1]].tolist() for u in range(g["num_users"])}
    return UniEvaluator(train, test, metric=METRIC_NAMES, top_k=list(top_k), batch_size=batch_size)


def test_full_gowalla_evaluation_matches_the_reference(gowalla):
    """All 29 858 test users x 40 981 items through UniEvaluator (tensor-core route: >= 16 384 items)
    against the reference's ProxyEvaluator string (np.matmul scores): every printed metric within
    1e-5 (north_star: NDCG@10 within 1e-5); a 512-user slice bit-exact against the C oracle; the SIMT
    and tensor-core paths bit-identical to each other on every user."""
    from neurec_b200 import ops
    g = gowalla
    U, V = gowalla_tables(g)
    dU, dV = dev(U), dev(V)

    class Model:
        def get_eval_tables(self):
            return dU, dV
    ev = _uni(g)
    assert ev.metrics_info() == g["kat"]["info"]
    text = ev.evaluate(Model())
    got, want = parse_result_string(text), parse_result_string(g["kat"]["eval_all_users"])
    assert got.shape == want.shape == (10,)
    assert np.abs(got - want).max() < 1e-5, (text, g["kat"]["eval_all_users"])
    assert abs(got[4] - want[4]) < 1e-5 and want[4] > 0.2             # NDCG@10, non-trivial
    sub = g["kat"]["subset_users"]
    text_sub = ev.evaluate(Model(), sub)
    assert np.abs(parse_result_string(text_sub) - parse_result_string(g["kat"]["eval_subset_512"])).max() < 1e-5

    metric = [1, 2, 4, 3, 5]
    users = np.arange(g["num_users"], dtype=np.int32)
    args = (dU, dV, dev(users), dev(g["train_indptr"]), dev(g["train_indices"]), dev(g["test_indptr"]),
            dev(g["test_indices"]), metric, 20)
    a = ops.eval_mf_tc(*args, return_ranks=True)
    b = ops.eval_mf(*args, return_ranks=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
    su = np.asarray(sub, dtype=np.int32)
    tp = np.zeros(len(su) + 1, np.int64); tp[1:] = np.cumsum(np.diff(g["test_indptr"])[su])
    ti = np.concatenate([g["test_indices"][g["test_indptr"][x]:g["test_indptr"][x + 1]] for x in su])
    rows, ranks = oracle.eval_mf(U, V, su, g["train_indptr"], g["train_indices"], tp, ti, metric, 20, thread_num=4,
                                 return_ranks=True)
    assert np.array_equal(a[1].cpu().numpy()[su], ranks) and np.array_equal(a[0].cpu().numpy()[su], rows)
