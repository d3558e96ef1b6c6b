"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference in this container.

Run from the repo root:  ``python tests/golden/make_golden.py``  (needs /root/reference and
``make -C oracle ref``).  The fixtures are committed; the GPU box never runs this script and
never reads /root/reference.

What is captured (all from the real reference code, imported through oracle.import_reference):
  kat_sampler.json    randint_choice / batch_randint_choice first outputs in a FRESH process
                      (glibc rand(), seed 1) + the first negatives of a PairwiseSampler epoch.
  kat_evaluator.npz   eval_score_matrix and arg_topk on random and tie-heavy score matrices.
  ml100k_split.npz    the ratio-0.8 split of ml-100k made by data.Dataset under
                      np.random.seed(2018) (train/test CSR, int16 item ids).
  kat_ml100k_eval.json  ProxyEvaluator.evaluate() strings on that split for random tables
                      (predict = np.matmul, exactly MF.py:120-122), topk=[10,20] and topk=5,
                      plus a user-subset (group) run.
  kat_surface.json    Configurator / DataIterator / sampler / metrics_info surface behaviour.
  gowalla_split.npz   (``python tests/golden/make_golden.py gowalla``) the 'given' split of
                      dataset/gowalla.{train,test} as loaded by data.Dataset (BASELINE config 3).
  ciao_split.npz / kat_ciao.json   (``python tests/golden/make_golden.py ciao``) dataset/Ciao_u5_s2 as loaded by
                      data.Dataset + SocialAbstractRecommender (trust CSR), SBPR._get_SocialItemsSet checksums and
                      4 000 (user, social item, negative, s_uk) samples of one real SBPR._get_pairwise_all_data epoch.
  kat_sampler_layout.json (``python tests/golden/make_golden.py layout``) one unshuffled epoch of the reference's PointwiseSampler
                      (neg_num=2) and PairwiseSampler (neg_num=3) on the ml-100k split: crc32 of users / labels / positives, shapes.
  kat_adjacency.json    (``python tests/golden/make_golden.py adjacency``) LightGCN.create_adj_mat for all five adj_type values and
                      NGCF.get_adj_mat('norm') run by the reference classes on the ml-100k split (nnz, crc32, sums).
  kat_neg_eval.json     (``python tests/golden/make_golden.py neg``) ProxyEvaluator.evaluate() of the reference with a negative-candidate
                      dict (rec.evaluate.neg > 0 branch, cpp/uni_evaluator.py:123-131) on the ml-100k split.
  kat_spectral.npz      (``python tests/golden/make_golden.py spectral``) SpectralCF's adjacency / degree / Laplacian methods
                      (SpectralCF.py:108-128) run by the real class on a 14 x 19 graph + the operator U U^T + U diag(lamda) U^T.
  kat_time_order.json   (same command) _generative_time_order_positive_items (data/sampler.py:42-68) run by the reference on the
                      by-time train sequences of that split, high_order 1..3: lengths and crc32 of its four outputs.
  kat_split_ml100k.npz  (``python tests/golden/make_golden.py split``) data/utils.py split_by_ratio(0.8) and split_by_loo with
                      by_time=True on ml-100k.rating: one train/test bit per interaction in file order.
  kat_gowalla.json / kat_gowalla_adj.npz   LightGCN.create_adj_mat('pre') (LightGCN.py:35-78) on that
                      split: nnz, row sums, value checksums; ProxyEvaluator strings for random
                      tables on all 29 858 test users and on a 512-user slice.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def fresh(code: str) -> str:
    """Run `code` in a fresh interpreter with the reference importable; return stdout."""
    pre = ("import sys, os, json; sys.path.insert(0, %r); import oracle; "
           "cwd = oracle.import_reference(); os.chdir(cwd); sys.argv=['main.py']\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", pre + code], capture_output=True, text=True,
                       check=True)
    return r.stdout


def kat_sampler():
    out = {}
    out["a"] = json.loads(fresh(
        "from util.cython.random_choice import randint_choice, batch_randint_choice\n"
        "a = randint_choice(100, size=5, exclusion=[1,2,3])\n"
        "b = randint_choice(40981, size=4)\n"
        "print(json.dumps({'a': list(a), 'b': list(b)}))").strip().splitlines()[-1])
    out["b"] = json.loads(fresh(
        "from util.cython.random_choice import randint_choice, batch_randint_choice\n"
        "a = randint_choice(100, size=5, exclusion=[1,2,3])\n"
        "c = batch_randint_choice(1682, [3,2], replace=True, exclusion=[[0,1],[5]])\n"
        "d = randint_choice(50, size=1, exclusion=list(range(40)))\n"
        "e = randint_choice(30, size=10, replace=False, exclusion=[0,1,2])\n"
        "print(json.dumps({'a': list(a), 'c': [list(x) for x in c], 'd': int(d), 'e': list(e)}))"
    ).strip().splitlines()[-1])
    # first negatives of a real PairwiseSampler epoch on the ml-100k split (fresh process)
    out["pairwise"] = json.loads(fresh(
        "import numpy as np, random\n"
        "np.random.seed(2018); random.seed(2018)\n"
        "from util import Configurator\n"
        "from data.dataset import Dataset\n"
        "from data import PairwiseSampler, PointwiseSampler\n"
        "conf = Configurator('NeuRec.properties', default_section='hyperparameters')\n"
        "ds = Dataset(conf)\n"
        "s = PairwiseSampler(ds, neg_num=1, batch_size=512, shuffle=False)\n"
        "it = iter(s); u, p, n = next(it)\n"
        "s2 = PointwiseSampler(ds, neg_num=2, batch_size=7, shuffle=False, drop_last=True)\n"
        "print(json.dumps({'len': len(s), 'users': u[:64], 'pos': [int(x) for x in p[:64]],"
        " 'neg': [int(x) for x in n[:64]], 'len_pointwise': len(s2)}))"
    ).strip().splitlines()[-1])
    with open(os.path.join(OUT, "kat_sampler.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    import oracle
    oracle.build()
    cwd = oracle.import_reference()
    os.chdir(cwd)
    sys.argv = ["main.py"]

    kat_sampler()

    from util.cython.arg_topk import arg_topk
    from evaluator.backend.cpp.cpp_evaluator import CPPEvaluator

    # ---- evaluator KATs -------------------------------------------------------------
    ev = CPPEvaluator()
    kat = {}
    S = np.random.RandomState(0).randn(4, 50).astype(np.float32)
    truth = [[0, 24, 43], [47], [5, 13] + list(range(30, 40)), [2]]
    kat["kat2_scores"] = S
    kat["kat2_truth_indptr"] = np.cumsum([0] + [len(t) for t in truth]).astype(np.int64)
    kat["kat2_truth_indices"] = np.concatenate([np.sort(t) for t in truth]).astype(np.int32)
    kat["kat2_out"] = ev.eval_score_matrix(S, truth, [1, 2, 3, 4, 5], 5, 1)
    kat["kat2_top5"] = arg_topk(S, 5, 1)
    # metric subsets / order as configured in NeuRec.properties:34 (P, R, NDCG, MAP, MRR)
    kat["kat2_out_41325"] = ev.eval_score_matrix(S, truth, [4, 1, 3, 2, 5], 5, 2)
    # ties (KAT-3 and more): integer-valued and -inf heavy rows
    z = np.zeros((1, 40), np.float32)
    kat["tie_zeros_top5"] = arg_topk(z, 5, 1)
    m3 = np.zeros((1, 40), np.float32); m3[0, ::3] = 1.0
    kat["tie_mult3_top8"] = arg_topk(m3, 8, 1)
    inf = np.full((1, 40), -np.inf, np.float32); inf[0, 7] = 1; inf[0, 3] = 2
    kat["tie_inf_top5"] = arg_topk(inf, 5, 1)
    rs = np.random.RandomState(7)
    T = rs.randint(0, 5, size=(64, 333)).astype(np.float32)
    T[rs.rand(64, 333) < 0.25] = -np.inf
    ttruth = [sorted(rs.choice(333, rs.randint(1, 25), replace=False).tolist()) for _ in range(64)]
    kat["tie_scores"] = T
    kat["tie_truth_indptr"] = np.cumsum([0] + [len(t) for t in ttruth]).astype(np.int64)
    kat["tie_truth_indices"] = np.concatenate(ttruth).astype(np.int32)
    kat["tie_out_k20"] = ev.eval_score_matrix(T.copy(), ttruth, [1, 2, 3, 4, 5], 20, 4)
    kat["tie_top20"] = arg_topk(T.copy(), 20, 4)
    kat["tie_top40"] = arg_topk(T.copy(), 40, 4)
    np.savez_compressed(os.path.join(OUT, "kat_evaluator.npz"), **kat)

    # ---- ml-100k split + full evaluator ---------------------------------------------
    np.random.seed(2018)
    import random
    random.seed(2018)
    from util import Configurator
    from data.dataset import Dataset
    from evaluator import ProxyEvaluator
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    ds = Dataset(conf)
    tr, te = ds.train_matrix.tocsr(), ds.test_matrix.tocsr()
    tr.sort_indices(); te.sort_indices()
    np.savez_compressed(os.path.join(OUT, "ml100k_split.npz"),
                        num_users=ds.num_users, num_items=ds.num_items,
                        train_indptr=tr.indptr.astype(np.int32), train_indices=tr.indices.astype(np.int16),
                        test_indptr=te.indptr.astype(np.int32), test_indices=te.indices.astype(np.int16))

    class _Model:
        def __init__(self, U, V):
            self.U, self.V = U, V

        def predict(self, user_ids, candidate_items=None):  # MF.py:120-124
            ratings = np.matmul(self.U[user_ids], self.V.T)
            if candidate_items is not None:
                ratings = [r[i] for r, i in zip(ratings, candidate_items)]
            return ratings

    rng = np.random.RandomState(1)
    U = (rng.randn(ds.num_users, 64) * .01).astype(np.float32)
    V = (rng.randn(ds.num_items, 64) * .01).astype(np.float32)
    model = _Model(U, V)
    res = {}
    train_d, test_d = ds.get_user_train_dict(), ds.get_user_test_dict()
    e1 = ProxyEvaluator(train_d, test_d, None, metric=conf["metric"], group_view=None,
                        top_k=conf["topk"], batch_size=conf["test_batch_size"], num_thread=8)
    res["info_topk_10_20"] = e1.metrics_info()
    res["eval_topk_10_20"] = e1.evaluate(model)
    e2 = ProxyEvaluator(train_d, test_d, None, metric=["NDCG", "Recall"], top_k=5, batch_size=100)
    res["info_topk_5"] = e2.metrics_info()
    res["eval_topk_5"] = e2.evaluate(model)
    # NOTE: GroupedEvaluator (grouped_evaluator.py:70-77) does not run under pandas 3.x
    # (groupby(by=[...]) yields tuple keys -> TypeError at :76), so no grouped golden exists;
    # its UniEvaluator-per-user-subset semantics (grouped_evaluator.py:108-110) is pinned by
    # evaluating explicit user subsets instead.
    sub = [u for u in test_d if 20 < len(train_d[u]) <= 50]
    res["subset_users"] = [int(u) for u in sub]
    res["eval_subset_20_50"] = e1.evaluator.evaluate(model, sub)
    # trained-looking tables (larger scale, correlated) for a second data point
    rng = np.random.RandomState(5)
    U2 = (rng.randn(ds.num_users, 32)).astype(np.float32)
    V2 = (rng.randn(ds.num_items, 32) + 0.3 * rng.randn(1, 32)).astype(np.float32)
    res["eval_topk_10_20_d32"] = e1.evaluate(_Model(U2, V2))
    res["dataset_str"] = str(ds)
    res["n_train"] = int(tr.nnz); res["n_test"] = int(te.nnz)
    with open(os.path.join(OUT, "kat_ml100k_eval.json"), "w") as f:
        json.dump(res, f, indent=1)

    # ---- surface behaviour ----------------------------------------------------------
    from util import DataIterator
    surf = {}
    surf["conf_recommender"] = conf["recommender"]
    surf["conf_topk"] = conf["topk"]
    surf["conf_sep"] = conf["data.convert.separator"]
    surf["conf_lr"] = conf["learning_rate"]
    surf["conf_params_str"] = conf.params_str()
    surf["conf_str"] = str(conf)
    di = DataIterator(list(range(10)), list(range(10, 20)), batch_size=4, shuffle=False)
    surf["dataiter"] = [b for b in di]
    surf["dataiter_len_drop"] = len(DataIterator(list(range(10)), batch_size=4, drop_last=True))
    np.random.seed(3)
    surf["dataiter_shuffle"] = [b for b in DataIterator(list(range(10)), batch_size=4, shuffle=True)]
    with open(os.path.join(OUT, "kat_surface.json"), "w") as f:
        json.dump(surf, f, indent=1)
    print("golden fixtures written to", OUT)


def gowalla():
    """BASELINE configs[2]: LightGCN on gowalla with the reference's own loader and adjacency."""
    import importlib
    import zlib
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    sys.argv = ["main.py", "--recommender=LightGCN", "--data.input.dataset=gowalla", "--splitter=given",
                "--data.column.format=UI", "--data.convert.separator=','", "--n_layers=3"]
    np.random.seed(2018)
    from util import Configurator
    from data.dataset import Dataset
    from evaluator import ProxyEvaluator
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    ds = Dataset(conf)
    tr, te = ds.train_matrix.tocsr(), ds.test_matrix.tocsr()
    tr.sort_indices(); te.sort_indices()
    assert ds.num_items < 65536
    np.savez_compressed(os.path.join(OUT, "gowalla_split.npz"), num_users=ds.num_users, num_items=ds.num_items,
                        train_indptr=tr.indptr.astype(np.int32), train_indices=tr.indices.astype(np.uint16),
                        test_indptr=te.indptr.astype(np.int32), test_indices=te.indices.astype(np.uint16))
    m = importlib.import_module("model.general_recommender.LightGCN")

    class _Self:
        pass
    f = _Self(); f.dataset = ds; f.n_users = ds.num_users; f.n_items = ds.num_items
    A = m.LightGCN.create_adj_mat(f, "pre").tocsr()
    A.sort_indices()
    np.savez_compressed(os.path.join(OUT, "kat_gowalla_adj.npz"),
                        rowsum=np.asarray(A.sum(1)).ravel().astype(np.float32),
                        row_nnz=np.diff(A.indptr).astype(np.int32),
                        data_head=A.data[:256].astype(np.float32), data_tail=A.data[-256:].astype(np.float32))
    res = {"num_users": int(ds.num_users), "num_items": int(ds.num_items), "train_nnz": int(tr.nnz),
           "test_nnz": int(te.nnz), "adj_shape": list(A.shape), "adj_nnz": int(A.nnz), "adj_dtype": str(A.dtype),
           "adj_sum_f64": float(A.data.astype(np.float64).sum()),
           "adj_indices_crc32": int(zlib.crc32(A.indices.astype(np.int32).tobytes())),
           "adj_data_crc32": int(zlib.crc32(A.data.astype(np.float32).tobytes())),
           "dataset_str": str(ds), "conf_n_layers": conf["n_layers"], "conf_batch_size": conf["batch_size"],
           "conf_lr": conf["lr"], "conf_reg": conf["reg"], "conf_embed_size": conf["embed_size"],
           "conf_adj_type": conf["adj_type"]}

    class _Model:
        def __init__(self, U, V):
            self.U, self.V = U, V

        def predict(self, user_ids, candidate_items=None):  # MF.py:120-124
            return np.matmul(self.U[user_ids], self.V.T)
    rng = np.random.RandomState(11)
    V = (rng.randn(ds.num_items, 64) * .1).astype(np.float32)
    train_d, test_d = ds.get_user_train_dict(), ds.get_user_test_dict()
    # trained-looking user rows: noise + the mean of a few of the user's held-out items, so that the
    # metric rows are non-trivial (hits at many ranks) -- same recipe in tests/test_gpu_gowalla.py
    U = (rng.randn(ds.num_users, 64) * .1).astype(np.float32)
    for u, items in test_d.items():
        U[u] += V[np.asarray(items[:3])].mean(0) * np.float32(1.5)
    ev = ProxyEvaluator(train_d, test_d, None, metric=conf["metric"], group_view=None, top_k=conf["topk"],
                        batch_size=conf["test_batch_size"], num_thread=8)
    res["info"] = ev.metrics_info()
    res["eval_all_users"] = ev.evaluate(_Model(U, V))
    sub = sorted(test_d.keys())[1000:1512]
    res["subset_users"] = [int(u) for u in sub]
    res["eval_subset_512"] = ev.evaluator.evaluate(_Model(U, V), sub)
    with open(os.path.join(OUT, "kat_gowalla.json"), "w") as fo:
        json.dump(res, fo, indent=1)
    print("gowalla fixtures written to", OUT)


def ciao():
    """SBPR on dataset/Ciao_u5_s2 with the reference's own loader, social matrix
    (SocialAbstractRecommender), social-item sets (SBPR._get_SocialItemsSet) and one real epoch of
    SBPR._get_pairwise_all_data (glibc rand() + np.random streams: the VALUES are not reproducible by the product,
    the (user, social item) -> s_uk relation and the membership constraints are)."""
    import importlib
    import zlib
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    sys.argv = ["main.py", "--recommender=SBPR", "--data.input.dataset=Ciao_u5_s2"]
    np.random.seed(2018)
    from util import Configurator
    from data.dataset import Dataset
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    ds = Dataset(conf)
    m = importlib.import_module("model.social_recommender.SBPR")
    model = m.SBPR(None, ds, conf)
    tr, te = ds.train_matrix.tocsr(), ds.test_matrix.tocsr()
    tr.sort_indices(); te.sort_indices()
    trust = model.social_matrix.tocsr(); trust.sort_indices()
    np.savez_compressed(os.path.join(OUT, "ciao_split.npz"), num_users=ds.num_users, num_items=ds.num_items,
                        train_indptr=tr.indptr.astype(np.int32), train_indices=tr.indices.astype(np.int32),
                        test_indptr=te.indptr.astype(np.int32), test_indices=te.indices.astype(np.int32),
                        trust_indptr=trust.indptr.astype(np.int32), trust_indices=trust.indices.astype(np.int32))
    social = model.userSocialItemsSetList
    sptr = np.zeros(ds.num_users + 1, np.int64)
    for u, items in social.items():
        sptr[u + 1] = len(items)
    sptr = np.cumsum(sptr)
    sidx = np.concatenate([np.sort(np.asarray(social[u], np.int32)) for u in sorted(social)]) if social else np.zeros(0, np.int32)
    users, pos, soc, neg, suk = (np.asarray(a) for a in model._get_pairwise_all_data())
    pick = np.random.RandomState(0).choice(len(users), 4000, replace=False)
    res = {"num_users": int(ds.num_users), "num_items": int(ds.num_items), "train_nnz": int(tr.nnz), "test_nnz": int(te.nnz),
           "trust_nnz": int(trust.nnz), "users_with_social_items": len(social), "social_items_total": int(sptr[-1]),
           "social_indptr_crc32": int(zlib.crc32(sptr.astype(np.int64).tobytes())),
           "social_indices_crc32": int(zlib.crc32(sidx.astype(np.int32).tobytes())),
           "epoch_samples": int(len(users)), "epoch_users_crc32": int(zlib.crc32(users.astype(np.int32).tobytes())),
           "epoch_pos_crc32": int(zlib.crc32(pos.astype(np.int32).tobytes())),
           "suk_mean": float(np.mean(suk)), "suk_max": int(np.max(suk)),
           "suk_hist": np.bincount(suk.astype(np.int64), minlength=8)[:8].tolist(),
           "sample_user": users[pick].astype(int).tolist(), "sample_social": soc[pick].astype(int).tolist(),
           "sample_neg": neg[pick].astype(int).tolist(), "sample_suk": suk[pick].astype(int).tolist(),
           "dataset_str": str(ds), "conf": {k: conf[k] for k in ("learning_rate", "embedding_size", "learner", "loss_function",
                                                                  "num_epochs", "reg_mf", "batch_size", "init_method", "stddev")}}
    with open(os.path.join(OUT, "kat_ciao.json"), "w") as fo:
        json.dump(res, fo, indent=1)
    print("ciao fixtures written to", OUT)


def by_time_dict(users, items, times, train_flags):
    """{dense user id: train items ordered by (time, file position)} -- shared by make_golden.py and tests/test_extras.py."""
    uid = np.unique(users, return_inverse=True)[1]
    keep = np.nonzero(train_flags)[0]
    order = keep[np.lexsort((keep, times[keep], uid[keep]))]
    out = {}
    for e in order:
        out.setdefault(int(uid[e]), []).append(int(items[e]))
    return out


def time_order(data, train_flags):
    """data/sampler.py:42-68 run by the REAL reference on the by-time train sequences of the split above."""
    import zlib
    from data.sampler import _generative_time_order_positive_items as gen
    d = by_time_dict(data["user"].values, data["item"].values, data["time"].values, train_flags)
    res = {}
    for ho in (1, 2, 3):
        lens, users, recent, nxt = gen(d, high_order=ho)
        res[str(ho)] = {"n": len(users), "lens_crc32": int(zlib.crc32(np.asarray(lens, np.int64).tobytes())),
                        "users_crc32": int(zlib.crc32(np.asarray(users, np.int32).tobytes())),
                        "recent_crc32": int(zlib.crc32(np.asarray(recent, np.int32).tobytes())),
                        "next_crc32": int(zlib.crc32(np.asarray(nxt, np.int32).tobytes()))}
    with open(os.path.join(OUT, "kat_time_order.json"), "w") as fo:
        json.dump(res, fo, indent=1)


def neg_eval():
    """rec.evaluate.neg > 0 (evaluator/backend/cpp/uni_evaluator.py:123-131): the REAL reference's ProxyEvaluator with a
    negative-candidate dict on the ml-100k split; inputs by the recipe of tests/test_surface.py::test_candidate_ranking_branch."""
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    from evaluator import ProxyEvaluator
    z = np.load(os.path.join(OUT, "ml100k_split.npz"))
    nu, ni = int(z["num_users"]), int(z["num_items"])
    rows = lambda p, i: {u: z[i][z[p][u]:z[p][u + 1]].astype(int).tolist() for u in range(nu) if z[p][u + 1] > z[p][u]}
    train_d, test_d = rows("train_indptr", "train_indices"), rows("test_indptr", "test_indices")
    rs = np.random.RandomState(2)
    neg_d = {}
    for u in test_d:
        seen = set(train_d[u]) | set(test_d[u])
        neg_d[u] = [int(i) for i in rs.choice(ni, 60) if i not in seen][:40]
    U = (rs.randn(nu, 16)).astype(np.float32); V = (rs.randn(ni, 16)).astype(np.float32)

    class _Model:
        def predict(self, user_ids, candidate_items=None):           # MF.py:120-134
            if candidate_items is None:
                return np.matmul(U[user_ids], V.T)
            return [np.squeeze(np.matmul(U[u], V[np.asarray(c)].T)) for u, c in zip(user_ids, candidate_items)]
    ev = ProxyEvaluator(train_d, test_d, neg_d, metric=["Recall", "NDCG"], group_view=None, top_k=10, batch_size=300, num_thread=8)
    res = {"info": ev.metrics_info(), "eval": ev.evaluate(_Model())}
    with open(os.path.join(OUT, "kat_neg_eval.json"), "w") as fo:
        json.dump(res, fo, indent=1)
    print("neg-eval fixture written:", res["eval"][:60])


def adjacency():
    """LightGCN.create_adj_mat for every adj_type (LightGCN.py:35-78) and NGCF.get_adj_mat('norm') (NGCF.py:288-319) run by
    the REAL reference classes on the ml-100k split: nnz, crc32 of the sorted-CSR index / fp32 value arrays, fp64 sum."""
    import importlib
    import types
    import zlib
    import scipy.sparse as sp
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    z = np.load(os.path.join(OUT, "ml100k_split.npz"))
    nu, ni = int(z["num_users"]), int(z["num_items"])
    users = np.repeat(np.arange(nu), np.diff(z["train_indptr"])).tolist()
    items = z["train_indices"].astype(int).tolist()

    def digest(A):
        A = A.tocoo().astype(np.float32).tocsr()
        A.sort_indices()
        return {"nnz": int(A.nnz), "indptr_crc32": int(zlib.crc32(A.indptr.astype(np.int64).tobytes())),
                "indices_crc32": int(zlib.crc32(A.indices.astype(np.int32).tobytes())),
                "data_crc32": int(zlib.crc32(A.data.astype(np.float32).tobytes())), "sum_f64": float(A.data.astype(np.float64).sum())}
    res = {}
    L = importlib.import_module("model.general_recommender.LightGCN").LightGCN
    ds = types.SimpleNamespace(get_train_interactions=lambda: (users, items))
    f = types.SimpleNamespace(dataset=ds, n_users=nu, n_items=ni)
    for t in ("plain", "norm", "gcmc", "pre", "mean"):
        res["lightgcn_" + t] = digest(L.create_adj_mat(f, t))
    N = importlib.import_module("model.general_recommender.NGCF").NGCF
    graph = sp.csr_matrix((np.ones(len(users), np.float32), (users, items)), shape=(nu, ni)).toarray()
    g = types.SimpleNamespace(num_users=nu, num_items=ni, graph=graph, adj_type="norm", logger=types.SimpleNamespace(info=lambda *a: None))
    g.normalized_adj_single = lambda adj: N.normalized_adj_single(g, adj)
    res["ngcf_norm"] = digest(N.get_adj_mat(g))
    with open(os.path.join(OUT, "kat_adjacency.json"), "w") as fo:
        json.dump(res, fo, indent=1)
    print("adjacency fixture written:", {k: v["nnz"] for k, v in res.items()})


def sampler_layout():
    """One unshuffled epoch of the REAL PointwiseSampler (neg_num=2) and PairwiseSampler (neg_num=3) on the ml-100k split
    (data/sampler.py:93-213): the layout -- who sits where -- as crc32s; the negatives themselves are glibc rand() draws."""
    import zlib
    out = json.loads(fresh(
        "import numpy as np, random, zlib\n"
        "np.random.seed(2018); random.seed(2018)\n"
        "from util import Configurator\n"
        "from data.dataset import Dataset\n"
        "from data import PairwiseSampler, PointwiseSampler\n"
        "conf = Configurator('NeuRec.properties', default_section='hyperparameters')\n"
        "ds = Dataset(conf)\n"
        "train = ds.get_user_train_dict()\n"
        "crc = lambda a, t: int(zlib.crc32(np.asarray(a, dtype=t).tobytes()))\n"
        "U, I, L = [], [], []\n"
        "for u, i, l in PointwiseSampler(ds, neg_num=2, batch_size=4096, shuffle=False): U += u; I += i; L += l\n"
        "n_pos = sum(len(v) for v in train.values())\n"
        "ok = all(I[e] not in train[U[e]] for e in range(n_pos, len(U)))\n"
        "res = {'n': len(U), 'n_pos': n_pos, 'users_crc32': crc(U, np.int32), 'labels_crc32': crc(L, np.float32),\n"
        "       'pos_items_crc32': crc(I[:n_pos], np.int32), 'negatives_outside_train': bool(ok)}\n"
        "PU, PP, PN = [], [], []\n"
        "for u, p, n in PairwiseSampler(ds, neg_num=3, batch_size=4096, shuffle=False): PU += u; PP += p; PN += n\n"
        "res.update({'pair_n': len(PU), 'pair_users_crc32': crc(PU, np.int32), 'pair_pos_crc32': crc(PP, np.int32),\n"
        "            'pair_neg_shape': list(np.asarray(PN).shape),\n"
        "            'pair_negatives_outside_train': bool(all(j not in train[u] for u, row in zip(PU, PN) for j in row))})\n"
        "print(json.dumps(res))").strip().splitlines()[-1])
    with open(os.path.join(OUT, "kat_sampler_layout.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("sampler layout fixture written:", out)


def spectral():
    """SpectralCF.adjacient_matrix / degree_matrix / laplacian_matrix (SpectralCF.py:108-128) run by the REAL reference class
    on a small bipartite graph (the methods need only self.graph / num_users / num_items), then the operator of :41-42,67-69."""
    import importlib
    import types
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    m = importlib.import_module("model.general_recommender.SpectralCF").SpectralCF
    rs = np.random.RandomState(7)
    nu, ni = 14, 19
    graph = (rs.rand(nu, ni) < 0.2).astype(np.float32)
    graph[np.arange(nu), rs.randint(0, ni, nu)] = 1.0               # no empty user
    f = types.SimpleNamespace(graph=graph, num_users=nu, num_items=ni)
    f.A = m.adjacient_matrix(f, self_connection=True)
    f.D = m.degree_matrix(f)
    L = m.laplacian_matrix(f, normalized=True)
    lamda, U = np.linalg.eig(L)
    lam = np.diag(lamda)
    A_hat = (np.dot(U, U.T) + np.dot(np.dot(U, lam), U.T)).astype(np.float32)       # SpectralCF.py:67-69
    np.savez_compressed(os.path.join(OUT, "kat_spectral.npz"), graph=graph.astype(np.uint8), A=f.A.astype(np.float32),
                        D=f.D.astype(np.float32), L=L.astype(np.float32), A_hat=A_hat)
    print("spectral fixture written", A_hat.shape, A_hat.dtype)


def split():
    """the reference's own split_by_ratio / split_by_loo (data/utils.py:59-106) with by_time=True
    on dataset/ml-100k.rating -> one train/test bit per interaction in FILE order (bit-packed)."""
    import oracle
    cwd = oracle.import_reference()
    os.chdir(cwd)
    import pandas as pd
    from data.utils import load_data, split_by_loo, split_by_ratio
    path = os.path.join("/root/reference", "dataset", "ml-100k.rating")
    cols = ["user", "item", "rating", "time"]
    out = {}
    for name, fn in (("ratio", lambda d: split_by_ratio(d, ratio=0.8, by_time=True)), ("loo", lambda d: split_by_loo(d, by_time=True))):
        data = load_data(path, "\t", cols)
        key = data["user"].astype(np.int64) * 1000003 + data["item"].astype(np.int64)
        assert key.is_unique
        train, test = fn(data.copy())
        tkey = set((train["user"].astype(np.int64) * 1000003 + train["item"].astype(np.int64)).tolist())
        flags = np.fromiter((k in tkey for k in key.tolist()), dtype=np.uint8, count=len(key))
        assert len(train) + len(test) == len(data) and int(flags.sum()) == len(train)
        out[name] = np.packbits(flags)
        out[name + "_train"] = np.int64(len(train))
    data = load_data(path, "\t", cols)                       # the inputs of the split, in file order (raw user ids, times)
    np.savez_compressed(os.path.join(OUT, "kat_split_ml100k.npz"), n=np.int64(len(data)),
                        user=data["user"].values.astype(np.uint16), item=data["item"].values.astype(np.uint16),
                        time=data["time"].values.astype(np.int32), **out)
    time_order(data, np.unpackbits(out["ratio"])[:len(data)])
    print("split fixtures written to", OUT, {k: int(v) for k, v in out.items() if k.endswith("_train")})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        split()
    elif len(sys.argv) > 1 and sys.argv[1] == "spectral":
        spectral()
    elif len(sys.argv) > 1 and sys.argv[1] == "neg":
        neg_eval()
    elif len(sys.argv) > 1 and sys.argv[1] == "adjacency":
        adjacency()
    elif len(sys.argv) > 1 and sys.argv[1] == "layout":
        sampler_layout()
    elif len(sys.argv) > 1 and sys.argv[1] == "gowalla":
        gowalla()
    elif len(sys.argv) > 1 and sys.argv[1] == "ciao":
        ciao()
    else:
        main()
