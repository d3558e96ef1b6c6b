"""The reference's plug-in surface (Configurator / DataIterator / samplers / ProxyEvaluator /
AbstractRecommender / main.py) on top of the kernels.  CPU tests pin the pure-Python pieces
against tests/golden/kat_surface.json (captured from the real reference); GPU tests drive the
whole stack."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, ROOT, parse_result_string


@pytest.fixture(scope="module")
def surf():
    with open(os.path.join(GOLDEN, "kat_surface.json")) as f:
        return json.load(f)


REF_MF_CONF = """[hyperparameters]
epochs=300
batch_size=512
embedding_size=64
reg_mf=0.0
learning_rate=0.001
learner=adam
num_negatives=1
#pairwise:BPR(BPRMF),hinge,square
is_pairwise=True
loss_function=bpr
init_method=normal
stddev=0.01
verbose=1"""


def test_configurator_matches_reference(surf, tmp_path, monkeypatch):
    from neurec_b200.util import Configurator
    (tmp_path / "conf").mkdir()
    (tmp_path / "conf" / "MF.properties").write_text(REF_MF_CONF)
    lib = open(os.path.join(ROOT, "NeuRec.properties")).read()
    (tmp_path / "NeuRec.properties").write_text(lib)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["main.py"])
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    assert conf["recommender"] == surf["conf_recommender"] == "MF"
    assert conf["topk"] == surf["conf_topk"] and conf["data.convert.separator"] == surf["conf_sep"]
    assert conf["learning_rate"] == surf["conf_lr"]
    assert conf.params_str() == surf["conf_params_str"]
    assert conf["is_pairwise"] is True and conf["group_view"] is None and conf["metric"][2] == "NDCG"
    assert "epochs" in conf and "nope" not in conf
    with pytest.raises(KeyError):
        conf["nope"]
    with pytest.raises(TypeError):
        conf[3]
    # command line: overrides keys present in a file, CLI-only keys are a last-resort lookup
    monkeypatch.setattr(sys, "argv", ["main.py", "--epochs=7", "--recommender=MF", "--my_flag=[1,2]"])
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    assert conf["epochs"] == 7 and conf["my_flag"] == [1, 2]
    monkeypatch.setattr(sys, "argv", ["main.py", "epochs=7"])
    with pytest.raises(SyntaxError):
        Configurator("NeuRec.properties")
    with pytest.raises(FileNotFoundError):
        Configurator("missing.properties")


def test_data_iterator_matches_reference(surf):
    from neurec_b200.util import DataIterator
    di = DataIterator(list(range(10)), list(range(10, 20)), batch_size=4, shuffle=False)
    assert [b for b in di] == surf["dataiter"] and len(di) == 3
    assert len(DataIterator(list(range(10)), batch_size=4, drop_last=True)) == surf["dataiter_len_drop"] == 2
    np.random.seed(3)
    assert [b for b in DataIterator(list(range(10)), batch_size=4, shuffle=True)] == surf["dataiter_shuffle"]
    with pytest.raises(ValueError):
        DataIterator([1, 2], [1], batch_size=2)
    with pytest.raises(ValueError):
        DataIterator([1, 2], batch_size=0)


def test_tool_helpers():
    from neurec_b200.util import csr_to_user_dict, pad_sequences, typeassert
    m = sp.csr_matrix(np.array([[0, 1, 1], [0, 0, 0], [1, 0, 1]]))
    assert csr_to_user_dict(m) == {0: [1, 2], 2: [0, 2]}            # empty rows omitted, ascending
    assert pad_sequences([[1, 2, 3], [4]], value=-1).tolist() == [[1, 2, 3], [4, -1, -1]]

    @typeassert(a=dict)
    def f(a):
        return a
    with pytest.raises(TypeError):
        f([1])


# ------------------------------------------------------------------------------------- GPU
def _ml100k_dataset(ml100k):
    from neurec_b200.data import Dataset
    d = ml100k
    shape = (d["num_users"], d["num_items"])
    tr = sp.csr_matrix((np.ones(len(d["train_indices"]), np.float32), d["train_indices"], d["train_indptr"]), shape=shape)
    te = sp.csr_matrix((np.ones(len(d["test_indices"]), np.float32), d["test_indices"], d["test_indptr"]), shape=shape)
    return Dataset.from_csr("ml-100k", tr, te)


@pytest.mark.gpu
def test_samplers_contract(ml100k, golden_sampler):
    from neurec_b200.data import PairwiseSampler, PointwiseSampler
    ds = _ml100k_dataset(ml100k)
    train = ds.get_user_train_dict()
    s = PairwiseSampler(ds, neg_num=1, batch_size=512, shuffle=False)
    assert len(s) == golden_sampler["pairwise"]["len"] == 157
    batches = list(s)
    assert len(batches) == 157 and len(batches[0][0]) == 512 and len(batches[-1][0]) == 80367 - 156 * 512
    u, p, n = batches[0]
    assert u[:64] == golden_sampler["pairwise"]["users"] and p[:64] == golden_sampler["pairwise"]["pos"]
    assert all(isinstance(x, int) for x in n) and all(nn not in train[uu] for uu, nn in zip(u, n))
    n2 = list(s)[0][2]
    assert n2 != n                                        # fresh negatives every __iter__
    import oracle
    sh = PairwiseSampler(ds, batch_size=1000, shuffle=True)
    a = list(sh)[0][0]
    flat_users = np.repeat(np.arange(ds.num_users), np.diff(ml100k["train_indptr"]))
    b = flat_users[oracle.shuffle_perm(80367, sh.seed, sh.epoch)[:1000]]
    assert a == b.tolist()                                # one keyed permutation per epoch (csrc/epoch.cuh)
    e0 = sh.epoch
    a2 = list(sh)[0][0]
    assert sh.epoch > e0 and a2 != a                      # a new order every __iter__
    # a NEW sampler continues the process-wide stream (MLP.py:100 builds one per epoch): fresh negatives
    n_a = list(PairwiseSampler(ds, batch_size=512, shuffle=False))[0][2]
    n_b = list(PairwiseSampler(ds, batch_size=512, shuffle=False))[0][2]
    assert n_a != n_b
    s3 = PairwiseSampler(ds, neg_num=3, batch_size=100, shuffle=False, drop_last=True)
    assert len(s3) == 803 and np.asarray(next(iter(s3))[2]).shape == (100, 3)
    pw = PointwiseSampler(ds, neg_num=2, batch_size=7, shuffle=False, drop_last=True)
    assert len(pw) == golden_sampler["pairwise"]["len_pointwise"]
    bu, bi, bl = next(iter(PointwiseSampler(ds, neg_num=2, batch_size=80367 * 3, shuffle=False)))
    assert bl[:80367] == [1.0] * 80367 and bl[80367:] == [0.0] * (2 * 80367)
    assert bu[:80367] == bu[80367:2 * 80367] and bi[:80367] == ml100k["train_indices"].tolist()
    assert all(i not in train[u] for u, i in zip(bu[80367:90000], bi[80367:90000]))
    with pytest.raises(ValueError):
        PairwiseSampler(ds, neg_num=0)


@pytest.mark.gpu
def test_randint_choice_api():
    from neurec_b200.util.random_choice import batch_randint_choice, randint_choice
    r = randint_choice(100, size=5, exclusion=[1, 2, 3])
    assert isinstance(r, list) and len(r) == 5 and not set(r) & {1, 2, 3}
    assert isinstance(randint_choice(50, size=1, exclusion=list(range(40))), int)
    e = randint_choice(30, size=10, replace=False, exclusion=[0, 1, 2])
    assert len(set(e)) == 10 and min(e) >= 3
    rows = batch_randint_choice(1682, [3, 1], replace=True, exclusion=[[0, 1], [5]])
    assert len(rows[0]) == 3 and isinstance(rows[1], int)
    for bad, exc in [(dict(high=5, size=0), ValueError), (dict(high=5, size=2, replace=1), TypeError),
                     (dict(high=5, size=2, p=[.2] * 5), NotImplementedError),
                     (dict(high=3, size=1, exclusion=[0, 1, 2]), ValueError),
                     (dict(high=5, size=4, replace=False, exclusion=[0]), ValueError)]:
        with pytest.raises(exc):
            randint_choice(**bad)
    with pytest.raises(ValueError):
        batch_randint_choice(10, [1, 2], exclusion=[[1]])


class _NumpyModel:
    """MF.predict of the reference (np.matmul), no fast path -> generic evaluator flow."""

    def __init__(self, U, V):
        self.U, self.V = U, V

    def predict(self, user_ids, candidate_items=None):
        ratings = np.matmul(self.U[user_ids], self.V.T)
        if candidate_items is not None:
            ratings = [r[i] for r, i in zip(ratings, candidate_items)]
        return ratings


@pytest.mark.gpu
def test_proxy_evaluator_reproduces_reference_strings(ml100k, golden_ml100k_eval):
    import torch
    from neurec_b200.evaluator import ProxyEvaluator
    g = golden_ml100k_eval
    ds = _ml100k_dataset(ml100k)
    train_d, test_d = ds.get_user_train_dict(), ds.get_user_test_dict()
    rng = np.random.RandomState(1)
    U = (rng.randn(ds.num_users, 64) * .01).astype(np.float32)
    V = (rng.randn(ds.num_items, 64) * .01).astype(np.float32)
    model = _NumpyModel(U, V)
    metric = ["Precision", "Recall", "NDCG", "MAP", "MRR"]
    e1 = ProxyEvaluator(train_d, test_d, None, metric=metric, group_view=None, top_k=[10, 20],
                        batch_size=128, num_thread=8)
    assert e1.metrics_info() == g["info_topk_10_20"]
    assert e1.evaluate(model) == g["eval_topk_10_20"]                 # byte-identical
    e2 = ProxyEvaluator(train_d, test_d, None, metric=["NDCG", "Recall"], top_k=5, batch_size=100)
    assert e2.metrics_info() == g["info_topk_5"] and e2.evaluate(model) == g["eval_topk_5"]
    assert e1.evaluator.evaluate(model, g["subset_users"]) == g["eval_subset_20_50"]
    rng = np.random.RandomState(5)
    U2 = rng.randn(ds.num_users, 32).astype(np.float32)
    V2 = (rng.randn(ds.num_items, 32) + 0.3 * rng.randn(1, 32)).astype(np.float32)
    assert e1.evaluate(_NumpyModel(U2, V2)) == g["eval_topk_10_20_d32"]

    class Fused(_NumpyModel):                                         # MF / LightGCN fast path
        def get_eval_tables(self):
            return torch.from_numpy(self.U).cuda(), torch.from_numpy(self.V).cuda()
    got = parse_result_string(e1.evaluate(Fused(U, V)))
    assert np.abs(got - parse_result_string(g["eval_topk_10_20"])).max() < 1e-5
    # grouped view = the same evaluator over user subsets (grouped_evaluator.py:63-112)
    eg = ProxyEvaluator(train_d, test_d, None, metric=metric, group_view=[20, 50, 100, 300], top_k=[10, 20],
                        batch_size=128)
    lines = eg.evaluate(model).split("\n")
    assert lines[0] == "" and [ln.split("\t")[0].strip() for ln in lines[1:]] == \
        ["(0,20]:", "(20,50]:", "(50,100]:", "(100,300]:"]
    assert lines[2].split("\t", 1)[1] == g["eval_subset_20_50"]
    with pytest.raises(ValueError):
        ProxyEvaluator(train_d, test_d, metric=["HitRatio"])
    with pytest.raises(TypeError):
        ProxyEvaluator([1], test_d)


@pytest.mark.gpu
def test_candidate_ranking_branch(ml100k):
    """rec.evaluate.neg > 0 (uni_evaluator.py:123-131) vs the oracle on the padded matrix."""
    import oracle
    from neurec_b200.evaluator import ProxyEvaluator
    ds = _ml100k_dataset(ml100k)
    train_d, test_d = ds.get_user_train_dict(), ds.get_user_test_dict()
    rs = np.random.RandomState(2)
    neg_d = {}
    for u in test_d:
        seen = set(train_d[u]) | set(test_d[u])
        neg_d[u] = [int(i) for i in rs.choice(ds.num_items, 60) if i not in seen][:40]
    U = (rs.randn(ds.num_users, 16)).astype(np.float32); V = (rs.randn(ds.num_items, 16)).astype(np.float32)
    model = _NumpyModel(U, V)
    ev = ProxyEvaluator(train_d, test_d, neg_d, metric=["Recall", "NDCG"], top_k=10, batch_size=300)
    got = parse_result_string(ev.evaluate(model))
    rows = []
    for u in test_d:
        c = list(test_d[u]) + neg_d[u]
        s = np.matmul(U[u], V[c].T)[None, :].astype(np.float32)
        pad = np.full((1, max(10, len(c))), -np.inf, np.float32); pad[0, :len(c)] = s
        ip, ix = oracle.lists_to_csr([range(len(test_d[u]))])
        rows.append(oracle.evaluate_matrix(pad, ip, ix, [2, 4], 10)[0])
    want = np.mean(np.stack(rows), axis=0)
    assert np.abs(got - want).max() < 1e-6


def _write_synthetic_dataset(path, nu=120, ni=200, seed=0):
    rs = np.random.RandomState(seed)
    lat_u, lat_i = rs.randn(nu, 4), rs.randn(ni, 4)
    rows = []
    for u in range(nu):
        p = np.exp(lat_u[u] @ lat_i.T); p /= p.sum()
        for i in rs.choice(ni, 25, replace=False, p=p):
            rows.append("%d\t%d\t%d\t%d" % (u + 1, i + 1, rs.randint(1, 6), 880000000 + rs.randint(10 ** 6)))
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "toy.rating"), "w") as f:
        f.write("\n".join(rows) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ["--recommender=MF", "--epochs=12", "--learning_rate=0.01"],
    ["--recommender=MF", "--epochs=3", "--is_pairwise=False", "--loss_function=cross_entropy", "--num_negatives=2"],
    ["--recommender=NeuMF", "--epochs=3", "--embedding_size=32"],
    ["--recommender=MLP", "--epochs=2"],
    ["--recommender=LightGCN", "--epochs=3", "--n_layers=3"],
    ["--recommender=NGCF", "--epochs=3", "--learning_rate=0.01"],
])
def test_main_end_to_end(tmp_path, args):
    """main.py + NeuRec.properties + conf/*.properties drive the kernels; the log lines keep the
    reference's format ("metrics:", "[iter e : loss : x, time: t]", "epoch e:\\t<values>")."""
    data = tmp_path / "dataset"
    _write_synthetic_dataset(str(data))
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--data.input.path=%s" % data,
           "--data.input.dataset=toy", "--topk=[5,10]", "--test_batch_size=64"] + args
    for name in ("NeuRec.properties", "conf"):
        os.symlink(os.path.join(ROOT, name), tmp_path / name)
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = r.stdout
    assert "metrics:\tPrecision@5 " in out and "NDCG@10" in out
    epochs = re.findall(r"epoch (\d+):\t([0-9.\t ]+)", out)
    assert len(epochs) >= 2
    vals = np.array([[float(x) for x in e[1].split()] for e in epochs])
    assert vals.shape[1] == 10 and np.isfinite(vals).all() and (vals >= 0).all() and (vals <= 1).all()
    if "LightGCN" not in " ".join(args):
        losses = [float(x) for x in re.findall(r"\[iter \d+ : loss : ([0-9.eE+-]+), time: ", out)]
        assert len(losses) == len(epochs) and losses[-1] < losses[0]       # it learns
    if args[0] == "--recommender=MF" and "--epochs=12" in args:
        assert vals[-1, 4] > vals[0, 4] * 1.2                               # NDCG@5 improves
    assert os.path.isdir(tmp_path / "log" / "toy")
    assert os.path.isfile(data / "_tmp_toy" / "toy_ratio_u0_i0.train")      # split cache


def test_product_adjacency_builder_equals_the_oracle(ml100k):
    """neurec_b200's own create_adj_mat restatement (used by the LightGCN plug-in and by bench.py)
    against the oracle's, for every adj_type of LightGCN.py:35-78 -- CPU only, no kernels."""
    from neurec_b200.model.general_recommender.LightGCN import bipartite_adjacency
    from oracle import tf_math
    d = ml100k
    nu, ni = d["num_users"], d["num_items"]
    users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    for adj_type in ("plain", "norm", "gcmc", "pre", "mean"):
        got = bipartite_adjacency(users, d["train_indices"], nu, ni, adj_type, verbose=False)
        got = got.tocoo().astype(np.float32).tocsr(); got.sort_indices()
        want = tf_math.lightgcn_adj(d["train_indptr"], d["train_indices"], nu, ni, adj_type)
        assert got.shape == want.shape and np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data), adj_type


def test_abstract_recommender_wires_evaluator_and_logger(ml100k, tmp_path, monkeypatch):
    """AbstractRecommender.__init__ (AbstractRecommender.py:23-36): ProxyEvaluator built from the
    dataset's three dicts and the evaluation options of NeuRec.properties, a log file under
    log/<dataset>/<model>/, dataset + configuration logged.  CPU only (no kernel is launched)."""
    import scipy.sparse as sp
    from neurec_b200.data import Dataset
    from neurec_b200.evaluator import ProxyEvaluator
    from neurec_b200.model.AbstractRecommender import AbstractRecommender
    from neurec_b200.util import Configurator
    (tmp_path / "conf").mkdir()
    (tmp_path / "conf" / "MF.properties").write_text(REF_MF_CONF)
    (tmp_path / "NeuRec.properties").write_text(open(os.path.join(ROOT, "NeuRec.properties")).read())
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["main.py"])
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    d = ml100k
    shape = (d["num_users"], d["num_items"])
    mk = lambda p, i: sp.csr_matrix((np.ones(len(d[i]), np.float32), d[i], d[p]), shape=shape)
    ds = Dataset.from_csr("ml-100k", mk("train_indptr", "train_indices"), mk("test_indptr", "test_indices"))
    model = AbstractRecommender(ds, conf)
    assert isinstance(model.evaluator, ProxyEvaluator)
    assert model.evaluator.metrics_info().startswith("metrics:\t")
    for k in conf["topk"]:
        assert "NDCG@%d" % k in model.evaluator.metrics_info()
    logs = list((tmp_path / "log" / "ml-100k" / "MF").glob("ml-100k_*.log"))
    assert len(logs) == 1 and "recommender" in logs[0].read_text()
    with pytest.raises(NotImplementedError):
        model.build_graph()
    with pytest.raises(NotImplementedError):
        model.predict([0], None)


REF_DATA = "/root/reference/dataset"


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF_DATA, "ml-100k.rating")), reason="the reference's dataset files are not here")
@pytest.mark.parametrize("name,golden", [("ml-100k", "ml100k_split.npz"), ("Ciao_u5_s2", "ciao_split.npz")])
def test_dataset_loader_reproduces_the_reference_split(tmp_path, monkeypatch, name, golden):
    """Our data.Dataset (load -> filter -> per-user ratio split under np.random.seed(2018) -> id remap -> CSR,
    data/dataset.py:75-210 + data/utils.py:24-80) against the train / test matrices the REAL reference's Dataset built
    from the same file (tests/golden/*_split.npz); for Ciao also SocialAbstractRecommender's trust matrix
    (AbstractRecommender.py:54-74).  CPU only; runs where the reference's dataset directory exists."""
    import shutil
    from neurec_b200.data import Dataset
    from neurec_b200.util import Configurator
    data = tmp_path / "dataset"
    data.mkdir()
    shutil.copy(os.path.join(REF_DATA, name + ".rating"), data / (name + ".rating"))
    for f in ("NeuRec.properties", "conf"):
        os.symlink(os.path.join(ROOT, f), tmp_path / f)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["main.py", "--data.input.path=%s" % data, "--data.input.dataset=%s" % name])
    np.random.seed(2018)                                   # main.py:10, as in make_golden.py
    conf = Configurator("NeuRec.properties", default_section="hyperparameters")
    ds = Dataset(conf)
    z = np.load(os.path.join(GOLDEN, golden))
    assert (ds.num_users, ds.num_items) == (int(z["num_users"]), int(z["num_items"]))
    for mat, p, i in ((ds.train_matrix, "train_indptr", "train_indices"), (ds.test_matrix, "test_indptr", "test_indices")):
        m = mat.tocsr(); m.sort_indices()
        assert np.array_equal(m.indptr, z[p].astype(m.indptr.dtype)) and np.array_equal(m.indices, z[i].astype(m.indices.dtype)), (name, p)
    if name == "Ciao_u5_s2":
        from neurec_b200.model.AbstractRecommender import SocialAbstractRecommender
        shutil.copy(os.path.join(REF_DATA, name + ".uu"), data / (name + ".uu"))
        monkeypatch.setattr(sys, "argv", sys.argv + ["--recommender=SBPR", "--social_file=%s" % (data / (name + ".uu"))])
        conf = Configurator("NeuRec.properties", default_section="hyperparameters")
        model = SocialAbstractRecommender(ds, conf)
        t = model.social_matrix.tocsr(); t.sort_indices()
        assert np.array_equal(t.indptr, z["trust_indptr"].astype(t.indptr.dtype))
        assert np.array_equal(t.indices, z["trust_indices"].astype(t.indices.dtype))
