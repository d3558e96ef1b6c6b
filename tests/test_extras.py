"""SURVEY.md 8(f) ranks 3-4 on the CPU: the oracle's restatements of SBPR / APR / CSR build pinned on what the REAL
reference produced on dataset/Ciao_u5_s2 (tests/golden/make_golden.py ciao) and on finite differences, and the
host-side logic of the product (social-item sets, time-ordered instance generation).  No kernel is launched."""
import json
import os
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from oracle import tf_math

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ciao():
    z = np.load(os.path.join(GOLDEN, "ciao_split.npz"))
    with open(os.path.join(GOLDEN, "kat_ciao.json")) as f:
        kat = json.load(f)
    d = {k: z[k] for k in z.files}
    d["kat"] = kat
    d["num_users"], d["num_items"] = int(z["num_users"]), int(z["num_items"])
    for k in ("train_indptr", "test_indptr", "trust_indptr"):
        d[k] = d[k].astype(np.int64)
    return d


@pytest.fixture(scope="module")
def ciao_social(ciao):
    return oracle.social_items_csr(ciao["train_indptr"], ciao["train_indices"], ciao["trust_indptr"], ciao["trust_indices"])


def test_social_item_sets_equal_the_reference(ciao, ciao_social):
    """SBPR._get_SocialItemsSet (SBPR.py:39-49) run by the real reference: row pointers and sorted items, crc for crc;
    the product's sparse-product formulation gives the same CSR."""
    kat = ciao["kat"]
    sptr, sidx = ciao_social
    assert int(sptr[-1]) == kat["social_items_total"] and int((np.diff(sptr) > 0).sum()) == kat["users_with_social_items"]
    assert zlib.crc32(sptr.astype(np.int64).tobytes()) == kat["social_indptr_crc32"]
    assert zlib.crc32(sidx.astype(np.int32).tobytes()) == kat["social_indices_crc32"]
    from neurec_b200.model.social_recommender.SBPR import social_items_csr
    shape = (ciao["num_users"], ciao["num_items"])
    train = sp.csr_matrix((np.ones(len(ciao["train_indices"]), np.float32), ciao["train_indices"], ciao["train_indptr"]), shape=shape)
    trust = sp.csr_matrix((np.ones(len(ciao["trust_indices"]), np.float32), ciao["trust_indices"], ciao["trust_indptr"]),
                          shape=(shape[0], shape[0]))
    p_ptr, p_idx = social_items_csr(train, trust)
    assert np.array_equal(p_ptr, sptr) and np.array_equal(p_idx, sidx)


def _contains(ptr, idx, row, x):
    seg = idx[ptr[row]:ptr[row + 1]]
    k = np.searchsorted(seg, x)
    return k < len(seg) and seg[k] == x


def test_restated_sampling_relations_hold_on_the_reference_epoch(ciao, ciao_social):
    """4 000 (user, social item, negative, s_uk) samples of one REAL SBPR._get_pairwise_all_data epoch: the relations the
    restatement (oracle/neurec_oracle.c::orc_sbpr_sample) is built on hold for every one of them."""
    kat = ciao["kat"]
    tp, ti, fp, fi = ciao["train_indptr"], ciao["train_indices"], ciao["trust_indptr"], ciao["trust_indices"]
    sptr, sidx = ciao_social
    for u, k, j, s in zip(kat["sample_user"], kat["sample_social"], kat["sample_neg"], kat["sample_suk"]):
        assert _contains(sptr, sidx, u, k) and not _contains(tp, ti, u, k)          # social item: a friend's, not the user's
        assert not _contains(tp, ti, u, j) and not _contains(sptr, sidx, u, j)      # negative: outside both
        assert s == 1 + sum(_contains(tp, ti, f, k) for f in fi[fp[u]:fp[u + 1]])   # s_uk


def test_oracle_sbpr_epoch_follows_the_contract(ciao, ciao_social):
    """The restated epoch (product order + product draws) against the reference's: same samples before shuffling
    (users / positives crc), every sample once, the relations above for every draw, s_uk histogram of the same law."""
    kat = ciao["kat"]
    tp, ti, fp, fi = ciao["train_indptr"], ciao["train_indices"], ciao["trust_indptr"], ciao["trust_indices"]
    sptr, sidx = ciao_social
    eligible = np.diff(sptr) > 0
    deg = np.diff(tp)
    pos_users = np.repeat(np.arange(ciao["num_users"], dtype=np.int32), np.where(eligible, deg, 0))
    pos_items = ti[np.repeat(eligible, deg)].astype(np.int32)
    assert len(pos_users) == kat["epoch_samples"]
    assert zlib.crc32(pos_users.astype(np.int32).tobytes()) == kat["epoch_users_crc32"]
    assert zlib.crc32(pos_items.astype(np.int32).tobytes()) == kat["epoch_pos_crc32"]
    ni = ciao["num_items"]
    u0, i0, k0, j0, s0 = oracle.sbpr_epoch_build(tp, ti, sptr, sidx, fp, fi, pos_users, pos_items, ni, False, 2018, 3)
    assert np.array_equal(u0, pos_users) and np.array_equal(i0, pos_items)
    u1, i1, k1, j1, s1 = oracle.sbpr_epoch_build(tp, ti, sptr, sidx, fp, fi, pos_users, pos_items, ni, True, 2018, 3)
    perm = oracle.shuffle_perm(len(pos_users), 2018, 3)
    assert np.array_equal(u1, u0[perm]) and np.array_equal(k1, k0[perm]) and np.array_equal(j1, j0[perm]) and np.array_equal(s1, s0[perm])
    u2, _, k2, j2, _ = oracle.sbpr_epoch_build(tp, ti, sptr, sidx, fp, fi, pos_users, pos_items, ni, False, 2018, 4)
    assert (k2 != k0).mean() > 0.5 and (j2 != j0).mean() > 0.99                    # a new epoch draws anew
    rs = np.random.RandomState(1)
    for q in rs.choice(len(u0), 3000, replace=False):
        u, k, j = int(u0[q]), int(k0[q]), int(j0[q])
        assert _contains(sptr, sidx, u, k) and not _contains(tp, ti, u, j) and not _contains(sptr, sidx, u, j)
        assert s0[q] == 1 + sum(_contains(tp, ti, f, k) for f in fi[fp[u]:fp[u + 1]])
    hist = np.bincount(s0.astype(np.int64), minlength=8)[:8]
    ref = np.asarray(kat["suk_hist"], np.float64)
    assert np.abs(hist - ref).max() < 6 * np.sqrt(ref.max())                         # same law, different stream
    assert abs(float(s0.mean()) - kat["suk_mean"]) < 0.02
    # the social item is uniform over the user's social row: chi^2 of the position quantile over all samples
    rel = np.array([np.searchsorted(sidx[sptr[u]:sptr[u + 1]], k) / max(sptr[u + 1] - sptr[u], 1)
                    for u, k in zip(u0[:20000], k0[:20000])])
    cnt = np.bincount(np.minimum((rel * 10).astype(int), 9), minlength=10)
    assert ((cnt - 2000.0) ** 2 / 2000.0).sum() < 60.0


@pytest.mark.parametrize("loss", ["bpr", "hinge", "square"])
def test_sbpr_grad_matches_finite_differences(loss):
    rs = np.random.RandomState(0)
    U = rs.randn(6, 5); V = rs.randn(9, 5); B = rs.randn(9)
    users = np.array([0, 1, 1, 5, 0]); pos = np.array([2, 2, 3, 6, 2]); soc = np.array([7, 8, 2, 0, 1])
    neg = np.array([4, 0, 5, 1, 5]); suk = np.array([1., 2., 3., 1., 4.])
    reg = 0.05
    _, gU, gV, gB, _, _ = tf_math.sbpr_grad(U, V, B, users, pos, soc, neg, suk, loss, reg)

    def f64(Ux, Vx, Bx):
        x = lambda it: (Ux[users] * Vx[it]).sum(1) + Bx[it]
        l = {"bpr": lambda r: np.log1p(np.exp(-r)), "hinge": lambda r: np.maximum(r + 1, 0), "square": lambda r: (1 - r) ** 2}[loss]
        tot = l((x(pos) - x(soc)) / suk).sum() + l(x(soc) - x(neg)).sum()
        sq = (Ux[users] ** 2).sum() + sum((Vx[it] ** 2).sum() + (Bx[it] ** 2).sum() for it in (pos, soc, neg))
        return tot + reg * 0.5 * sq

    def num_grad(x):
        g = np.zeros_like(x)
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            k = it.multi_index
            old = x[k]
            x[k] = old + 1e-5; fp_ = f64(U, V, B)
            x[k] = old - 1e-5; fm = f64(U, V, B)
            x[k] = old
            g[k] = (fp_ - fm) / 2e-5
        return g
    for got, x in ((gU, U), (gV, V), (gB, B)):
        assert np.allclose(got, num_grad(x), rtol=2e-3, atol=3e-4)


def test_sbpr_trainer_learns_and_moves_every_variable():
    rs = np.random.RandomState(2)
    U = (rs.randn(30, 8) * .1).astype(np.float32); V = (rs.randn(50, 8) * .1).astype(np.float32)
    B = (rs.randn(50) * .1).astype(np.float32)
    users = rs.randint(0, 30, 400); pos = rs.randint(0, 25, 400); soc = rs.randint(25, 40, 400); neg = rs.randint(40, 50, 400)
    suk = rs.randint(1, 4, 400).astype(np.float32)
    tr = tf_math.SBPRTrainer(U, V, B, "adam", 0.01, "bpr", 0.01)
    first = tr.epoch(users, pos, soc, neg, suk, 64).sum()
    for _ in range(15):
        last = tr.epoch(users, pos, soc, neg, suk, 64).sum()
    assert last < 0.7 * first
    assert np.abs(tr.U - U).max() > 1e-3 and np.abs(tr.V - V).max() > 1e-3 and np.abs(tr.B - B).max() > 1e-3


def test_l2_normalize_rows_restatement():
    rs = np.random.RandomState(3)
    x = rs.randn(7, 16).astype(np.float32)
    x[2] = 0.0
    got = tf_math.l2_normalize_rows(x, 0.5)
    want = x.astype(np.float64) / np.sqrt(np.maximum((x.astype(np.float64) ** 2).sum(1, keepdims=True), 1e-12)) * 0.5
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7) and np.all(got[2] == 0)
    assert np.allclose(np.linalg.norm(got[[0, 1, 3]], axis=1), 0.5, rtol=1e-5)


def test_csr_from_coo_restatement_equals_scipy():
    rs = np.random.RandomState(4)
    rows = rs.randint(0, 40, 900); cols = rs.randint(0, 70, 900)
    ptr, idx = oracle.csr_from_coo(rows, cols, 40)
    m = sp.csr_matrix((np.ones(900), (rows, cols)), shape=(40, 70))
    m.sum_duplicates(); m.sort_indices()
    assert np.array_equal(ptr, m.indptr) and np.array_equal(idx, m.indices)


def test_time_order_instances_follow_the_reference_loop():
    """_generative_time_order_positive_items (data/sampler.py:42-68) restated with numpy windows against the loop as written."""
    from neurec_b200.data.sampler import _generative_time_order_positive_items as gen
    d = {0: [5, 3, 9, 1], 1: [2], 2: [7, 8], 4: [1, 2, 3, 4, 5, 6]}
    for ho in (1, 2, 3):
        lens, users, recent, nxt = gen(d, ho)
        w_len, w_u, w_r, w_n = [], [], [], []
        for u, seq in d.items():
            if len(seq) - ho <= 0:
                continue
            m = len(seq) - ho
            w_len.append([u, m]); w_u += [u] * m
            w_r += [seq[i] for i in range(m)] if ho == 1 else [seq[i:][:ho] for i in range(m)]
            w_n += seq[ho:]
        assert lens.tolist() == w_len and users.tolist() == w_u and recent.tolist() == w_r and nxt.tolist() == w_n
    with pytest.raises(ValueError):
        gen(d, 0)
    with pytest.raises(TypeError):
        gen([1, 2], 1)
    with pytest.raises(ValueError):
        gen({}, 1)


# ------------------------------------------------------------------------------------------------ SpectralCF
def _spectral_problem(dtype, seed=0, nu=9, ni=13, d=5, K=2):
    rs = np.random.RandomState(seed)
    rows = [np.sort(rs.choice(ni, rs.randint(1, 5), replace=False)) for _ in range(nu)]
    ptr = np.cumsum([0] + [len(r) for r in rows]); idx = np.concatenate(rows)
    A = tf_math.spectralcf_a_hat(ptr, idx, nu, ni).astype(dtype)
    e0 = (rs.randn(nu + ni, d) * 0.3).astype(dtype)
    W = [(rs.randn(d, d) * 0.4).astype(dtype) for _ in range(K)]
    users = rs.randint(0, nu, 7); pos = rs.randint(0, ni, 7); neg = rs.randint(0, ni, 7)
    return ptr, idx, A, e0, W, nu, users, pos, neg


@pytest.mark.parametrize("act", ["sigmoid", "tanh", "relu", "elu", "identity", "selu"])
def test_spectralcf_gradients_match_finite_differences(act):
    """Manual backprop of SpectralCF.py:63-91 (concat, activation, filter product, dense spectral product) against
    central differences in fp64 -- for every activation of util/tool.py:10-33 the kernels provide."""
    _, _, A, e0, W, nu, users, pos, neg = _spectral_problem(np.float64)
    total = lambda e, w: tf_math.spectralcf_loss_and_grad(A, e, w, nu, users, pos, neg, 0.05, "bpr", act)[0]
    _, dE0, dW, _ = tf_math.spectralcf_loss_and_grad(A, e0, W, nu, users, pos, neg, 0.05, "bpr", act)
    rs = np.random.RandomState(1)
    h = 1e-6
    for _ in range(12):
        r, c = rs.randint(e0.shape[0]), rs.randint(e0.shape[1])
        p, m = e0.copy(), e0.copy(); p[r, c] += h; m[r, c] -= h
        fd = (total(p, W) - total(m, W)) / (2 * h)
        assert abs(fd - dE0[r, c]) < 1e-6 * max(1.0, abs(fd))
    for k in range(len(W)):
        for _ in range(6):
            r, c = rs.randint(W[k].shape[0]), rs.randint(W[k].shape[1])
            Wp = [w.copy() for w in W]; Wm = [w.copy() for w in W]; Wp[k][r, c] += h; Wm[k][r, c] -= h
            fd = (total(e0, Wp) - total(e0, Wm)) / (2 * h)
            assert abs(fd - dW[k][r, c]) < 1e-6 * max(1.0, abs(fd))
    with pytest.raises(NotImplementedError):
        tf_math.activation("softplus", e0)


def test_spectral_operator_of_the_product_equals_the_restatement():
    """SpectralCF.__init__ (:37-43) + A_hat (:67-69): the plug-in's host-side construction and the oracle's, same bits;
    the operator is symmetric up to rounding and A_hat 1 = (1 + eigen-part) behaves like a smoothing operator."""
    from neurec_b200.model.general_recommender.SpectralCF import spectral_operator
    ptr, idx, A, *_ = _spectral_problem(np.float32, nu=14, ni=19)
    train = sp.csr_matrix((np.ones(len(idx), np.float32), idx, ptr), shape=(14, 19))
    got = spectral_operator(train)
    assert got.dtype == np.float32 and got.shape == (33, 33) and np.array_equal(got, A)
    assert np.abs(got - got.T).max() < 1e-5


def test_spectralcf_trainer_learns():
    ptr, idx, A, e0, W, nu, users, pos, neg = _spectral_problem(np.float32, seed=3, nu=20, ni=30, d=8)
    rs = np.random.RandomState(0)
    users = np.repeat(np.arange(nu), np.diff(ptr)); pos = idx
    tr = tf_math.SpectralCFTrainer(A, e0, W, nu, "adam", 0.01, 1e-3)
    losses = []
    for _ in range(40):
        neg = rs.randint(0, 30, len(users))
        losses.append(float(tr.step(users, pos, neg)))
    assert losses[-1] < 0.8 * losses[0]
    assert tr.embeddings().shape == (50, 24)


# ------------------------------------------------------------------------------------------ train / test split
@pytest.fixture(scope="module")
def split_golden():
    z = np.load(os.path.join(GOLDEN, "kat_split_ml100k.npz"))
    n = int(z["n"])
    users = np.unique(z["user"], return_inverse=True)[1].astype(np.int32)
    return {"n": n, "users": users, "num_users": int(users.max()) + 1, "time": z["time"].astype(np.int64),
            "ratio": np.unpackbits(z["ratio"])[:n], "loo": np.unpackbits(z["loo"])[:n]}


@pytest.mark.parametrize("mode", ["ratio", "loo"])
def test_split_restatement_equals_the_reference_split(split_golden, mode):
    """data/utils.py split_by_ratio(0.8) / split_by_loo with by_time=True run by the REAL reference on ml-100k.rating
    (100 000 interactions, many equal timestamps inside a user): the restatement assigns every interaction to the
    same side."""
    g = split_golden
    got = oracle.split_interactions(g["users"], g["time"], g["num_users"], mode, 0.8)
    assert np.array_equal(got, g[mode])


def test_random_split_restatement_contract(split_golden):
    """by_time=False: a uniformly random ceil(ratio n_u) of every user's interactions, another one for another seed."""
    g = split_golden
    a = oracle.split_interactions(g["users"], None, g["num_users"], "ratio", 0.8, seed=1)
    b = oracle.split_interactions(g["users"], None, g["num_users"], "ratio", 0.8, seed=2)
    cnt = np.bincount(g["users"], minlength=g["num_users"])
    want = np.ceil(0.8 * cnt).astype(np.int64)
    assert np.array_equal(np.bincount(g["users"], weights=a, minlength=g["num_users"]).astype(np.int64), want)
    assert np.array_equal(np.bincount(g["users"], weights=b, minlength=g["num_users"]).astype(np.int64), want)
    assert 0.25 < (a != b).mean() < 0.40                       # two independent 80 % subsets differ on ~32 %
    # position inside the user's time order does not matter: early and late interactions are kept equally often
    order = np.lexsort((g["time"], g["users"]))
    first_half = np.zeros(g["n"], bool)
    start = 0
    for c in cnt:
        first_half[order[start:start + c // 2]] = True
        start += c
    assert abs(a[first_half].mean() - a[~first_half].mean()) < 0.01


def test_time_order_instances_equal_the_reference_on_ml100k(split_golden):
    """_generative_time_order_positive_items run by the REAL reference (tests/golden/make_golden.py split ->
    kat_time_order.json) on the by-time train sequences of the ratio-0.8 split of ml-100k: our numpy-window
    restatement yields the same lengths, users, windows and next items (crc32) for high_order 1, 2, 3."""
    from neurec_b200.data.sampler import _generative_time_order_positive_items as gen
    z = np.load(os.path.join(GOLDEN, "kat_split_ml100k.npz"))
    with open(os.path.join(GOLDEN, "kat_time_order.json")) as f:
        kat = json.load(f)
    users, items, times = z["user"].astype(np.int64), z["item"].astype(np.int64), z["time"].astype(np.int64)
    flags = split_golden["ratio"]
    uid = np.unique(users, return_inverse=True)[1]                      # same construction as make_golden.by_time_dict
    keep = np.nonzero(flags)[0]
    order = keep[np.lexsort((keep, times[keep], uid[keep]))]
    d = {}
    for e in order:
        d.setdefault(int(uid[e]), []).append(int(items[e]))
    for ho in (1, 2, 3):
        lens, us, recent, nxt = gen(d, high_order=ho)
        want = kat[str(ho)]
        assert len(us) == want["n"]
        assert zlib.crc32(np.asarray(lens, np.int64).tobytes()) == want["lens_crc32"]
        assert zlib.crc32(np.asarray(us, np.int32).tobytes()) == want["users_crc32"]
        assert zlib.crc32(np.ascontiguousarray(recent, dtype=np.int32).tobytes()) == want["recent_crc32"]
        assert zlib.crc32(np.asarray(nxt, np.int32).tobytes()) == want["next_crc32"]


def test_spectral_operator_equals_the_reference_methods():
    """SpectralCF.adjacient_matrix / degree_matrix / laplacian_matrix of the REAL reference class on a 14 x 19 graph
    (kat_spectral.npz) and the operator built from their eigendecomposition: the oracle's restatement and the plug-in's
    host-side construction reproduce A, D, L exactly and A_hat to the rounding of numpy's eig (same LAPACK here: equal)."""
    from neurec_b200.model.general_recommender.SpectralCF import spectral_operator
    z = np.load(os.path.join(GOLDEN, "kat_spectral.npz"))
    graph = z["graph"].astype(np.float32)
    nu, ni = graph.shape
    train = sp.csr_matrix(graph)
    got = spectral_operator(train)
    want = z["A_hat"]
    assert got.shape == want.shape == (nu + ni, nu + ni)
    assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
    ora = tf_math.spectralcf_a_hat(train.indptr, train.indices, nu, ni)
    assert np.abs(ora - want).max() < 1e-5 * np.abs(want).max()
    # the pieces, restated here the way both constructions build them
    A = np.identity(nu + ni, dtype=np.float32); A[:nu, nu:] += graph; A[nu:, :nu] += graph.T
    assert np.array_equal(A, z["A"]) and np.array_equal(A.sum(1), z["D"])
    L = np.identity(nu + ni, dtype=np.float32) - np.dot(np.diag(np.power(A.sum(1), -1)), A)
    assert np.array_equal(L, z["L"])
