"""GPU parity of the SURVEY.md 8(f) rank 3-4 kernels (csrc/extras.cu) through the C ABI against the oracle:
APR's row normaliser and epoch, SBPR's device-built epoch (bit-exact vs oracle/neurec_oracle.c::orc_sbpr_sample),
quadruple step and epoch loop (vs tf_math.SBPRTrainer), the time-ordered samplers, COO -> CSR, checkpoint/resume."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from oracle import tf_math

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ciao():
    z = np.load(os.path.join(GOLDEN, "ciao_split.npz"))
    d = {k: z[k] for k in z.files}
    d["num_users"], d["num_items"] = int(z["num_users"]), int(z["num_items"])
    for k in ("train_indptr", "test_indptr", "trust_indptr"):
        d[k] = d[k].astype(np.int64)
    for k in ("train_indices", "test_indices", "trust_indices"):
        d[k] = d[k].astype(np.int32)
    sptr, sidx = oracle.social_items_csr(d["train_indptr"], d["train_indices"], d["trust_indptr"], d["trust_indices"])
    d["social_indptr"], d["social_indices"] = sptr, sidx
    eligible = np.diff(sptr) > 0
    deg = np.diff(d["train_indptr"])
    d["pos_users"] = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.where(eligible, deg, 0))
    d["pos_items"] = d["train_indices"][np.repeat(eligible, deg)]
    d["max_excluded"] = int((deg + np.diff(sptr))[eligible].max())
    return d


# ------------------------------------------------------------------------------------------------ APR
def test_l2_normalize_rows_vs_oracle():
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    for rows, dim in ((943, 64), (1682, 64), (7, 16), (33, 100)):
        x = (rs.randn(rows, dim) * rs.rand(rows, 1) * 3).astype(np.float32)
        x[rows // 2] = 0.0
        got = ops.l2_normalize_rows(dev(x), 0.5).cpu().numpy()
        want = tf_math.l2_normalize_rows(x, 0.5)
        assert np.abs(got - want).max() < 5e-7 and np.all(got[rows // 2] == 0)
    xd = dev(x)
    ops.l2_normalize_rows(xd, 2.0, xd)                                   # in place
    assert np.abs(xd.cpu().numpy() - tf_math.l2_normalize_rows(x, 2.0)).max() < 1e-6


class _Conf(dict):
    def params_str(self):
        return "test"


def _dataset(d, name="ml-100k"):
    from neurec_b200.data import Dataset
    shape = (d["num_users"], d["num_items"])
    mk = lambda p, i: sp.csr_matrix((np.ones(len(d[i]), np.float32), d[i], d[p]), shape=shape)
    return Dataset.from_csr(name, mk("train_indptr", "train_indices"), mk("test_indptr", "test_indices"))


BASE_CONF = {"metric": ["Precision", "Recall", "NDCG", "MAP", "MRR"], "group_view": None, "topk": [10, 20],
             "test_batch_size": 128, "num_thread": 8, "data.convert.separator": "\t"}


def test_apr_epoch_is_the_bpr_softplus_epoch_and_adversarial_ops(ml100k, tmp_path, monkeypatch):
    """APR.train_model optimises self.loss (APR.py:120-122): one epoch through the plug-in equals tf_math.MFTrainer
    (bpr, reg 0) fed the oracle's restatement of the same device epoch; update_adversarial's deltas equal
    l2_normalize(grad) * eps of the oracle (APR.py:106-118)."""
    from neurec_b200.data import PairwiseSampler, sampler
    from neurec_b200.model.general_recommender.APR import APR
    monkeypatch.chdir(tmp_path)
    d = ml100k
    conf = _Conf(BASE_CONF, recommender="APR", learning_rate=1e-3, embedding_size=64, learner="adam", epochs=1, eps=0.5,
                 adv="grad", adver=1, adv_epoch=0, reg=0.0, reg_adv=1.0, batch_size=512, init_method="tnormal",
                 stddev=0.01, verbose=1)
    model = APR(None, _dataset(d), conf)
    model.build_graph()
    P0, Q0 = model.embedding_P.cpu().numpy().copy(), model.embedding_Q.cpu().numpy().copy()
    assert np.abs(P0).max() <= 0.02 + 1e-7                                 # truncated normal, 2 sigma
    sampler.reseed(40)
    it = PairwiseSampler(model.dataset, neg_num=1, batch_size=512, shuffle=True)
    total = model._train_epoch(it)
    users = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
    wu, wi, wj = oracle.epoch_build(d["train_indptr"], d["train_indices"], users, d["train_indices"], 1, d["num_items"],
                                    True, True, 2018, 40)
    tr = tf_math.MFTrainer(P0, Q0, "adam", 1e-3, "bpr", 0.0, True)
    want = tr.epoch(wu, wi, wj[:, 0], 512)
    assert abs(total - float(want.sum())) < 1e-4 * float(want.sum())
    assert np.abs(model.embedding_P.cpu().numpy() - tr.U).max() < 2e-5
    assert np.abs(model.embedding_Q.cpu().numpy() - tr.V).max() < 2e-5
    # adversarial ops: gradient of the plain BPR loss of one batch, rows normalised, times eps
    bu, bi, bj = wu[:512], wi[:512], wj[:512, 0]
    model.update_adversarial(bu, bi, bj)
    _, gU, gV, _, _ = tf_math.mf_pairwise_grad(model.embedding_P.cpu().numpy(), model.embedding_Q.cpu().numpy(), bu, bi, bj,
                                               "bpr", 0.0)
    # the kernel sums duplicate rows with atomics (another order than np.add.at): tolerance on the normalised rows
    assert np.abs(model.delta_P.cpu().numpy() - tf_math.l2_normalize_rows(gU, 0.5)).max() < 2e-5
    assert np.abs(model.delta_Q.cpu().numpy() - tf_math.l2_normalize_rows(gV, 0.5)).max() < 2e-5
    touched = np.zeros(d["num_users"], bool); touched[bu] = True
    norms = np.linalg.norm(model.delta_P.cpu().numpy(), axis=1)
    assert np.allclose(norms[touched], 0.5, atol=1e-4) and np.all(norms[~touched] == 0)
    assert float(model._gP.abs().max().item()) == 0.0                       # accumulators left clean
    model.adv = "random"
    model.update_adversarial()
    assert np.allclose(np.linalg.norm(model.delta_Q.cpu().numpy(), axis=1), 0.5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ SBPR
@pytest.mark.parametrize("shuffle,epoch", [(True, 0), (False, 7), (True, 123456789012)])
def test_sbpr_epoch_build_bit_exact(ciao, shuffle, epoch):
    from neurec_b200 import ops
    c = ciao
    args = [c[k] for k in ("train_indptr", "train_indices", "social_indptr", "social_indices", "trust_indptr", "trust_indices",
                           "pos_users", "pos_items")]
    want = oracle.sbpr_epoch_build(*args, c["num_items"], shuffle, 2018, epoch)
    got = ops.sbpr_epoch_build(*[dev(a) for a in args], c["num_items"], c["max_excluded"], shuffle, 2018, epoch)
    for g, w, name in zip(got, want, ("users", "pos", "social", "neg", "suk")):
        assert np.array_equal(g.cpu().numpy(), w), name
    # any window of the epoch is the same slice (an epoch does not depend on how it is cut)
    win = ops.sbpr_epoch_build(*[dev(a) for a in args], c["num_items"], c["max_excluded"], shuffle, 2018, epoch, 1000, 5000)
    for g, w in zip(win, want):
        assert np.array_equal(g.cpu().numpy(), w[1000:6000])
    with pytest.raises(ValueError):
        ops.sbpr_epoch_build(*[dev(a) for a in args], c["num_items"], c["num_items"], shuffle, 2018, epoch)


def _sbpr_problem(seed=0, nu=300, ni=500, dim=16, n=4096):
    rs = np.random.RandomState(seed)
    U = (rs.randn(nu, dim) * 0.1).astype(np.float32); V = (rs.randn(ni, dim) * 0.1).astype(np.float32)
    B = (rs.randn(ni) * 0.1).astype(np.float32)
    users = rs.randint(0, nu, n).astype(np.int32)
    pos = rs.randint(0, ni, n).astype(np.int32); soc = rs.randint(0, ni, n).astype(np.int32)
    neg = rs.randint(0, ni, n).astype(np.int32)
    suk = rs.randint(1, 6, n).astype(np.float32)
    return U, V, B, users, pos, soc, neg, suk


@pytest.mark.parametrize("loss", ["bpr", "hinge", "square"])
def test_sbpr_grad_vs_oracle(loss):
    from neurec_b200 import ops
    U, V, B, users, pos, soc, neg, suk = _sbpr_problem()
    want_l, gU, gV, gB, tU, tV = tf_math.sbpr_grad(U, V, B, users, pos, soc, neg, suk, loss, 0.01)
    dU, dV, dB = dev(U), dev(V), dev(B)
    z = torch.zeros_like
    hU, hV, hB = z(dU), z(dV), z(dB)
    sU = torch.zeros(U.shape[0], dtype=torch.int32, device="cuda"); sV = torch.zeros(V.shape[0], dtype=torch.int32, device="cuda")
    out = torch.zeros(1, device="cuda")
    ops.sbpr_grad(dU, dV, dB, dev(users), dev(pos), dev(soc), dev(neg), dev(suk), loss, 0.01, hU, hV, hB, sU, sV, 9, out)
    assert abs(out.item() - float(want_l)) < 1e-4 * abs(float(want_l))
    scale = max(np.abs(gU).max(), np.abs(gV).max())
    assert np.abs(hU.cpu().numpy() - gU).max() < 1e-5 * max(1.0, scale)
    assert np.abs(hV.cpu().numpy() - gV).max() < 1e-5 * max(1.0, scale)
    assert np.abs(hB.cpu().numpy() - gB).max() < 1e-5 * max(1.0, np.abs(gB).max())
    assert np.array_equal(sU.cpu().numpy() == 9, tU) and np.array_equal(sV.cpu().numpy() == 9, tV)
    with pytest.raises(ValueError):
        ops.sbpr_grad(dU, dV, dB, dev(users), dev(pos), dev(soc), dev(neg), dev(suk), "cross_entropy", 0.0, hU, hV, hB, sU, sV, 9, out)


@pytest.mark.parametrize("opt", ["adam", "gd", "adagrad", "rmsprop", "momentum"])
def test_sbpr_train_epoch_vs_oracle(opt):
    from neurec_b200 import ops
    U, V, B, users, pos, soc, neg, suk = _sbpr_problem(seed=1, n=512 * 6 - 77)
    lr = {"adam": 1e-3, "gd": 0.05, "adagrad": 0.01, "rmsprop": 1e-3, "momentum": 0.02}[opt]
    tr = tf_math.SBPRTrainer(U, V, B, opt, lr, "bpr", 0.01)
    want = tr.epoch(users, pos, soc, neg, suk, 512)
    steps = len(want)
    dU, dV, dB = dev(U), dev(V), dev(B)
    i0, i1 = tf_math.SLOT_INIT[opt]
    mk = lambda a, v: None if v is None else torch.full_like(a, v)
    slots = [(mk(t, i0), mk(t, i1)) for t in (dU, dV, dB)]
    z = torch.zeros_like
    tU = torch.zeros(U.shape[0], dtype=torch.int32, device="cuda"); tV = torch.zeros(V.shape[0], dtype=torch.int32, device="cuda")
    step_loss = torch.zeros(steps, device="cuda")
    lr_t = tf_math.adam_lr_t(lr, steps) if opt == "adam" else np.full(steps, lr, np.float32)
    n = ops.sbpr_train_epoch(dU, dV, dB, dev(users), dev(pos), dev(soc), dev(neg), dev(suk), 512, "bpr", 0.01, opt, lr_t,
                             tf_math.DEFAULT_HYPER[opt](lr), z(dU), z(dV), z(dB), tU, tV, slots[0][0], slots[0][1],
                             slots[1][0], slots[1][1], slots[2][0], slots[2][1], 1, step_loss)
    assert n == steps
    assert np.allclose(step_loss.cpu().numpy(), want, rtol=1e-4)
    for got, ref, name in ((dU, tr.U, "U"), (dV, tr.V, "V"), (dB, tr.B, "bias")):
        assert np.abs(got.cpu().numpy() - ref).max() < 3e-5, name
    assert np.abs(tr.B - B).max() > 1e-4 and np.abs(tr.U - U).max() > 1e-4


def _write_social_dataset(path, nu=150, ni=260, seed=0):
    rs = np.random.RandomState(seed)
    lat_u, lat_i = rs.randn(nu, 4), rs.randn(ni, 4)
    rows, pairs = [], []
    for u in range(nu):
        p = np.exp(lat_u[u] @ lat_i.T); p /= p.sum()
        for i in rs.choice(ni, 22, replace=False, p=p):
            rows.append("%d\t%d\t%d\t%d" % (u + 1, i + 1, rs.randint(1, 6), 880000000 + rs.randint(10 ** 6)))
        sim = lat_u @ lat_u[u]
        sim[u] = -1e9
        for f in np.argsort(-sim)[:rs.randint(1, 6)]:
            pairs.append("%d\t%d" % (u + 1, f + 1))
    pairs.append("%d\t%d" % (nu + 50, 1))                                 # an unknown user: dropped by the loader
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "toy.rating"), "w") as f:
        f.write("\n".join(rows) + "\n")
    with open(os.path.join(path, "toy.uu"), "w") as f:
        f.write("\n".join(pairs) + "\n")


@pytest.mark.parametrize("args", [["--recommender=SBPR", "--num_epochs=6", "--learning_rate=0.01"],
                                  ["--recommender=APR", "--epochs=6", "--learning_rate=0.01"]])
def test_main_runs_the_new_plug_ins(tmp_path, args):
    data = tmp_path / "dataset"
    _write_social_dataset(str(data))
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--data.input.path=%s" % data, "--data.input.dataset=toy",
           "--topk=[5,10]", "--test_batch_size=64", "--social_file=%s" % (data / "toy.uu")] + args
    for name in ("NeuRec.properties", "conf"):
        os.symlink(os.path.join(ROOT, name), tmp_path / name)
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = r.stdout
    assert "metrics:\tPrecision@5 " in out and "NDCG@10" in out
    epochs = re.findall(r"epoch (\d+):\t([0-9.\t ]+)", out)
    assert len(epochs) >= 5
    vals = np.array([[float(x) for x in e[1].split()] for e in epochs])
    assert vals.shape[1] == 10 and np.isfinite(vals).all() and (vals >= 0).all() and (vals <= 1).all()
    losses = [float(x) for x in re.findall(r"\[iter \d+ : loss : ([0-9.eE+-]+), time: ", out)]
    assert len(losses) == len(epochs) and losses[-1] < losses[0]


# ---------------------------------------------------------------------------------- time-ordered samplers
def _timed_dataset(nu=60, ni=90, seed=0):
    from neurec_b200.data import Dataset
    rs = np.random.RandomState(seed)
    rows, cols, times = [], [], []
    for u in range(nu):
        items = rs.choice(ni, rs.randint(1, 14), replace=False)
        rows += [u] * len(items); cols += items.tolist(); times += rs.permutation(len(items) * 3)[:len(items)].tolist()
    train = sp.csr_matrix((np.ones(len(rows), np.float32), (rows, cols)), shape=(nu, ni))
    tm = sp.csr_matrix((np.asarray(times, np.float32) + 1.0, (rows, cols)), shape=(nu, ni))
    test = sp.csr_matrix(([1.0], ([0], [0])), shape=(nu, ni))
    return Dataset.from_csr("toy", train, test, time_matrix=tm)


@pytest.mark.parametrize("high_order,neg_num", [(1, 1), (2, 3), (3, 2)])
def test_time_order_samplers_contract(high_order, neg_num):
    """data/sampler.py:216-354: every (user, window, next item) instance exactly once per epoch, its window intact
    after the shuffle, negatives never an item of the user, pointwise layout positives + k-major negatives."""
    from neurec_b200.data import TimeOrderPairwiseSampler, TimeOrderPointwiseSampler
    ds = _timed_dataset()
    seqs = ds.get_user_train_dict(by_time=True)
    truth = {}
    for u, s in seqs.items():
        for t in range(len(s) - high_order):
            truth[(u, s[t + high_order])] = s[t] if high_order == 1 else list(s[t:t + high_order])
    pw = TimeOrderPairwiseSampler(ds, high_order=high_order, neg_num=neg_num, batch_size=64, shuffle=True)
    assert len(pw) == (len(truth) + 63) // 64
    seen, order = set(), []
    for bu, br, bp, bn in pw:
        assert len(bu) <= 64 and len(bu) == len(br) == len(bp) == len(bn)
        for u, r, p, n in zip(bu, br, bp, bn):
            assert truth[(u, p)] == r
            for j in ([n] if neg_num == 1 else n):
                assert j not in seqs[u] and 0 <= j < ds.num_items
            assert (u, p) not in seen
            seen.add((u, p)); order.append((u, p))
    assert seen == set(truth)
    assert order != [o for b in pw for o in zip(b[0], b[2])]                # a new order (and new negatives) every epoch
    pt = TimeOrderPointwiseSampler(ds, high_order=high_order, neg_num=neg_num, batch_size=50, shuffle=True, drop_last=True)
    n_all = len(truth) * (neg_num + 1)
    assert len(pt) == n_all // 50 and len(pt.users_list) == n_all and len(pt.recent_items_list) == n_all
    pos_seen, n_neg = set(), 0
    for bu, br, bi, bl in pt:
        assert len(bu) == 50
        for u, r, i, l in zip(bu, br, bi, bl):
            if l == 1.0:
                assert truth[(u, i)] == r
                pos_seen.add((u, i))
            else:
                assert l == 0.0 and i not in seqs[u]
                n_neg += 1
    assert len(pos_seen) + n_neg == len(pt) * 50 and pos_seen <= set(truth)
    with pytest.raises(ValueError):
        TimeOrderPairwiseSampler(ds, high_order=0)
    with pytest.raises(ValueError):
        TimeOrderPointwiseSampler(ds, neg_num=0)


def test_gather_rows_i32():
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    src = rs.randint(0, 1000, (37, 3)).astype(np.int32)
    idx = rs.randint(0, 37 * 4, 500).astype(np.int64)
    assert np.array_equal(ops.gather_rows_i32(dev(src), dev(idx)).cpu().numpy(), src[idx % 37])
    assert np.array_equal(ops.gather_rows_i32(dev(src[:, 0].copy()), dev(idx)).cpu().numpy(), src[idx % 37, 0])


# ------------------------------------------------------------------------------------------ COO -> CSR
def test_csr_from_coo_vs_oracle(ml100k):
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    for nr, nc, nnz in ((40, 70, 900), (1000, 50, 3000), (5, 100000, 20000), (300, 300, 0)):
        rows = rs.randint(0, nr, nnz).astype(np.int32); cols = rs.randint(0, nc, nnz).astype(np.int32)
        if nnz:
            rows[rows == 3] = 4                                          # an empty row
        ptr, idx = ops.csr_from_coo(dev(rows), dev(cols), nr, nc)
        wp, wi = oracle.csr_from_coo(rows, cols, nr)
        assert np.array_equal(ptr.cpu().numpy(), wp) and np.array_equal(idx.cpu().numpy(), wi)
    d = ml100k                                                           # the real interactions, shuffled and with repeats
    users = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
    p = rs.permutation(len(users))
    rows = np.concatenate([users[p], users[p[:5000]]]); cols = np.concatenate([d["train_indices"][p], d["train_indices"][p[:5000]]])
    ptr, idx = ops.csr_from_coo(dev(rows), dev(cols), d["num_users"], d["num_items"])
    assert np.array_equal(ptr.cpu().numpy(), d["train_indptr"]) and np.array_equal(idx.cpu().numpy(), d["train_indices"])
    bad = rows.copy(); bad[17] = d["num_users"]
    with pytest.raises(ValueError):
        ops.csr_from_coo(dev(bad), dev(cols), d["num_users"], d["num_items"])


# ------------------------------------------------------------------------------------- checkpoint / resume
def test_checkpoint_resume_continues_the_run(ml100k, tmp_path, monkeypatch):
    """Two epochs, save, a third epoch  ==  load into a fresh model + the third epoch: tables, Adam state, beta
    powers, stamps and the sampler stream all come back (neurec_b200/util/checkpoint.py).  The restored state is
    bit-identical; the third epochs agree within the re-association of the gradient atomics (two runs of the SAME
    process differ by as much)."""
    from neurec_b200.data import PairwiseSampler, sampler
    from neurec_b200.model.general_recommender.MF import MF
    from neurec_b200.util import checkpoint
    monkeypatch.chdir(tmp_path)
    conf = _Conf(BASE_CONF, recommender="MF", learning_rate=1e-3, embedding_size=64, learner="adam", loss_function="bpr",
                 is_pairwise=True, epochs=3, reg_mf=0.001, batch_size=512, verbose=1, num_negatives=1,
                 init_method="normal", stddev=0.01)
    ds = _dataset(ml100k)

    def fresh():
        m = MF(None, ds, conf)
        m.build_graph()
        return m, PairwiseSampler(ds, neg_num=1, batch_size=512, shuffle=True)
    sampler.reseed(70)
    a, it = fresh()
    a._train_epoch(it); a._train_epoch(it)
    path = str(tmp_path / "mf.ckpt")
    checkpoint.save(a, path)
    la = a._train_epoch(it)
    sampler.reseed(0)                                                    # a new process would start anywhere
    b, it_b = fresh()
    meta = checkpoint.load(b, path)
    assert meta["model"] == "MF" and meta["sampler_stream"] == 72
    saved = torch.load(path, map_location="cpu")["tensors"]
    for name in ("user_embeddings", "item_embeddings", "_s0U", "_s1U", "_s0V", "_s1V", "_tU", "_tV"):
        assert torch.equal(getattr(b, name).cpu(), saved[name]), name       # restored bit for bit
    assert torch.equal(b.opt.device_pows().cpu(), saved["opt._pows_dev"]) and b.opt.stamp == a.opt.stamp - len(it)
    lb = b._train_epoch(it_b)
    assert abs(la - lb) < 1e-5 * abs(la)
    for name in ("user_embeddings", "item_embeddings", "_s0U", "_s1U", "_s0V", "_s1V"):
        assert (getattr(a, name) - getattr(b, name)).abs().max().item() < 2e-6, name
    assert (a.user_embeddings - dev(saved["user_embeddings"].numpy())).abs().max().item() > 1e-4   # the epoch moved them
    conf2 = _Conf(conf, learner="gd")
    c = MF(None, ds, conf2); c.build_graph()
    with pytest.raises(ValueError):
        checkpoint.load(c, path)


# ------------------------------------------------------------------------------------------------ SpectralCF
def _spectral_case(nu=120, ni=170, d=24, K=2, seed=0, batch=256):
    rs = np.random.RandomState(seed)
    rows = [np.sort(rs.choice(ni, rs.randint(2, 12), replace=False)) for _ in range(nu)]
    ptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64); idx = np.concatenate(rows).astype(np.int32)
    A = tf_math.spectralcf_a_hat(ptr, idx, nu, ni)
    e0 = (rs.randn(nu + ni, d) * 0.2).astype(np.float32)
    W = (rs.randn(K, d, d) * (1.0 / np.sqrt(d))).astype(np.float32)
    users = rs.randint(0, nu, batch).astype(np.int32)
    pos = rs.randint(0, ni, batch).astype(np.int32); neg = rs.randint(0, ni, batch).astype(np.int32)
    return A, e0, W, nu, users, pos, neg


@pytest.mark.parametrize("act,d,K", [("sigmoid", 24, 2), ("tanh", 100, 2), ("relu", 33, 1), ("elu", 16, 3), ("identity", 128, 1),
                                     ("selu", 8, 2)])
def test_spectralcf_forward_vs_oracle(act, d, K):
    from neurec_b200 import ops
    A, e0, W, nu, *_ = _spectral_case(d=d, K=K)
    want, _ = tf_math.spectralcf_forward(A, e0, list(W), act)
    got = ops.spectralcf_forward(dev(A), dev(e0), dev(W), act).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    with pytest.raises(NotImplementedError):
        ops.spectralcf_forward(dev(A), dev(e0), dev(W), "softmax")


@pytest.mark.parametrize("act,loss,with_t", [("sigmoid", "bpr", True), ("tanh", "hinge", False), ("relu", "square", True),
                                             ("selu", "BPR", False)])
def test_spectralcf_grad_vs_oracle(act, loss, with_t):
    """Loss, d loss / d E_0 and d loss / d W_k of one batch against the numpy backprop (itself checked by finite
    differences); tolerance: fp32 re-association of the dense products (fixed-order k-chunks here, BLAS blocking there)."""
    from neurec_b200 import ops
    A, e0, W, nu, users, pos, neg = _spectral_case(d=24, K=2, seed=1)
    N, d, K = e0.shape[0], e0.shape[1], W.shape[0]
    want_l, dE0, dW, want_all = tf_math.spectralcf_loss_and_grad(A, e0, list(W), nu, users, pos, neg, 0.01, loss, act)
    dA = dev(A)
    all_emb = torch.zeros((N, d * (K + 1)), device="cuda"); G = torch.zeros_like(all_emb)
    touched = torch.zeros(N, dtype=torch.int32, device="cuda")
    gE = torch.empty((N, d), device="cuda"); gW = torch.empty((K, d, d), device="cuda")
    out = torch.zeros(1, device="cuda")
    ops.spectralcf_grad(nu, dA, dA.t().contiguous() if with_t else None, dev(e0), dev(W), act, dev(users), dev(pos), dev(neg),
                        loss, 0.01, all_emb, G, touched, gE, gW, ops.spectralcf_work(N, d, K), out)
    assert abs(out.item() - float(want_l)) < 1e-4 * abs(float(want_l))
    assert np.abs(all_emb.cpu().numpy() - want_all).max() < 2e-5
    assert np.abs(gE.cpu().numpy() - dE0).max() < 2e-5 * max(1.0, np.abs(dE0).max())
    for k in range(K):
        assert np.abs(gW[k].cpu().numpy() - dW[k]).max() < 5e-5 * max(1.0, np.abs(dW[k]).max()), k
    assert float(G.abs().max().item()) == 0.0                            # the accumulator is handed back clean


def test_spectralcf_steps_vs_oracle_trainer():
    """Five Adam steps through nrc_spectralcf_grad + nrc_opt_apply_multi (dense formulas) against SpectralCFTrainer."""
    from neurec_b200 import ops
    A, e0, W, nu, *_ = _spectral_case(d=16, K=2, seed=2)
    N, d, K = e0.shape[0], 16, 2
    rs = np.random.RandomState(5)
    tr = tf_math.SpectralCFTrainer(A, e0, list(W), nu, "adam", 1e-2, 1e-3, "bpr", "sigmoid")
    dA, dE, dWt = dev(A), dev(e0), dev(W)
    dAt = dA.t().contiguous()
    all_emb = torch.zeros((N, d * (K + 1)), device="cuda"); G = torch.zeros_like(all_emb)
    touched = torch.zeros(N, dtype=torch.int32, device="cuda")
    gE, gW = torch.zeros_like(dE), torch.zeros_like(dWt)
    z = torch.zeros_like
    sE, sW = (z(dE), z(dE)), (z(dWt), z(dWt))
    work = ops.spectralcf_work(N, d, K)
    lr_t = tf_math.adam_lr_t(1e-2, 5)
    for s in range(5):
        users = rs.randint(0, nu, 200).astype(np.int32); pos = rs.randint(0, 170, 200).astype(np.int32)
        neg = rs.randint(0, 170, 200).astype(np.int32)
        want = tr.step(users, pos, neg)
        out = torch.zeros(1, device="cuda")
        ops.spectralcf_grad(nu, dA, dAt, dE, dWt, "sigmoid", dev(users), dev(pos), dev(neg), "bpr", 1e-3, all_emb, G, touched, gE, gW,
                            work, out)
        ops.opt_apply_multi("adam", [(dE, gE, sE[0], sE[1], None, True), (dWt, gW, sW[0], sW[1], None, True)], s + 1,
                            [float(lr_t[s]), 0.9, 0.999, 1e-8])
        assert abs(out.item() - float(want)) < 2e-4 * abs(float(want)), s
    assert np.abs(dE.cpu().numpy() - tr.e0).max() < 1e-4
    for k in range(K):
        assert np.abs(dWt[k].cpu().numpy() - tr.filters[k]).max() < 1e-4
    assert np.abs(tr.e0 - e0).max() > 1e-2                                # five Adam steps of 1e-2 moved the table


def test_main_runs_spectralcf(tmp_path):
    data = tmp_path / "dataset"
    _write_social_dataset(str(data))
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--data.input.path=%s" % data, "--data.input.dataset=toy",
           "--topk=[5,10]", "--test_batch_size=64", "--recommender=SpectralCF", "--epochs=5", "--embedding_size=32",
           "--learning_rate=0.005"]
    for name in ("NeuRec.properties", "conf"):
        os.symlink(os.path.join(ROOT, name), tmp_path / name)
    r = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    epochs = re.findall(r"epoch (\d+):\t([0-9.\t ]+)", r.stdout)
    assert len(epochs) == 5
    vals = np.array([[float(x) for x in e[1].split()] for e in epochs])
    assert vals.shape[1] == 10 and np.isfinite(vals).all() and (vals >= 0).all() and (vals <= 1).all()
    losses = [float(x) for x in re.findall(r"\[iter \d+ : loss : ([0-9.eE+-]+), time: ", r.stdout)]
    assert len(losses) == 5 and losses[-1] < losses[0]


# ------------------------------------------------------------------------------------------ train / test split
def test_split_interactions_bit_exact_and_equal_to_the_reference_split():
    """nrc_split_interactions against the restatement (every mode) and, for by_time=True on ml-100k, against the bits
    the REAL reference's split_by_ratio / split_by_loo produced (tests/golden/kat_split_ml100k.npz)."""
    from neurec_b200 import ops
    z = np.load(os.path.join(GOLDEN, "kat_split_ml100k.npz"))
    n = int(z["n"])
    users = np.unique(z["user"], return_inverse=True)[1].astype(np.int32)
    nu = int(users.max()) + 1
    times = z["time"].astype(np.int64)
    for mode in ("ratio", "loo"):
        got = ops.split_interactions(dev(users), dev(times), nu, mode, 0.8).cpu().numpy()
        assert np.array_equal(got, np.unpackbits(z[mode])[:n]), mode
        assert np.array_equal(got, oracle.split_interactions(users, times, nu, mode, 0.8))
    for seed in (0, 2018):
        for mode, ratio in (("ratio", 0.8), ("ratio", 0.5), ("loo", 0.0)):
            got = ops.split_interactions(dev(users), None, nu, mode, ratio, seed).cpu().numpy()
            assert np.array_equal(got, oracle.split_interactions(users, None, nu, mode, ratio, seed)), (seed, mode, ratio)
    rs = np.random.RandomState(0)                               # negative times, empty users, tiny users
    u = rs.randint(0, 40, 300).astype(np.int32); u[u == 7] = 8
    t = rs.randint(-50, 50, 300).astype(np.int64)
    for mode in ("ratio", "loo"):
        assert np.array_equal(ops.split_interactions(dev(u), dev(t), 40, mode, 0.7).cpu().numpy(),
                              oracle.split_interactions(u, t, 40, mode, 0.7))
    with pytest.raises(ValueError):
        ops.split_interactions(dev(u), dev(t), 39, "ratio", 0.8)
    with pytest.raises(ValueError):
        ops.split_interactions(dev(u), dev(t), 40, "given", 0.8)


def test_csr_row_ids(ml100k):
    from neurec_b200 import ops
    d = ml100k
    want = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
    assert np.array_equal(ops.csr_row_ids(dev(d["train_indptr"])).cpu().numpy(), want)
    ptr = np.array([0, 0, 3, 3, 3, 1000, 1001], np.int64)                 # empty rows, a long row
    assert np.array_equal(ops.csr_row_ids(dev(ptr)).cpu().numpy(), np.repeat(np.arange(6, dtype=np.int32), np.diff(ptr)))
