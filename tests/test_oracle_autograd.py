"""Independent check of every manual backprop in oracle/tf_math.py: the reference's graphs are written a second time as
torch (CPU, float64) forward passes straight from the reference's definitions, and torch.autograd's gradients are compared
with the oracle's hand-derived ones on random problems with duplicate ids.  (TensorFlow itself cannot be installed here;
this pins the DIFFERENTIATION -- the part of the restatement most likely to hide a slip -- on an independent
implementation; the forward definitions are checked against the cited reference lines by reading.)  CPU only."""
import numpy as np
import pytest
import torch

from oracle import tf_math

T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64, requires_grad=True)
I = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int64))


def _pair_loss(kind, x):          # util/learner.py:19-29
    if kind == "bpr":
        return -torch.nn.functional.logsigmoid(x).sum()
    if kind == "hinge":
        return torch.clamp(x + 1.0, min=0).sum()
    return ((1.0 - x) ** 2).sum()


def _point_loss(kind, z, x):      # util/learner.py:31-41 (tf.losses.sigmoid_cross_entropy = mean reduction)
    if kind == "cross_entropy":
        return torch.nn.functional.binary_cross_entropy_with_logits(x, z, reduction="mean")
    return ((z - x) ** 2).sum()


def _l2(*ts):                     # util/tool.py:216-217  sum(t^2) / 2
    return sum((t ** 2).sum() for t in ts) / 2


def _ids(rs, n, hi):
    a = rs.randint(0, hi, n)
    a[1] = a[0]                   # duplicates inside the batch
    return a


@pytest.mark.parametrize("loss", ["bpr", "hinge", "square"])
def test_mf_pairwise(loss):       # MF.py:54-68
    rs = np.random.RandomState(0)
    U, V = rs.randn(9, 6), rs.randn(11, 6)
    u, i, j = _ids(rs, 20, 9), _ids(rs, 20, 11), _ids(rs, 20, 11)
    l, gU, gV, _, _ = tf_math.mf_pairwise_grad(U.astype(np.float64), V.astype(np.float64), u, i, j, loss, 0.03)
    tU, tV = T(U), T(V)
    p, qi, qj = tU[I(u)], tV[I(i)], tV[I(j)]
    total = _pair_loss(loss, (p * qi).sum(1) - (p * qj).sum(1)) + 0.03 * _l2(p, qj, qi)
    total.backward()
    assert abs(float(total.detach()) - float(l)) < 1e-5 * abs(float(total.detach()))           # the oracle's fp32 loss
    assert np.allclose(gU, tU.grad.numpy(), rtol=2e-5, atol=2e-6) and np.allclose(gV, tV.grad.numpy(), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("pairwise,loss,n_towers", [(False, "cross_entropy", 1), (False, "square", 1), (True, "bpr", 2),
                                                    (True, "hinge", 2), (True, "bpr", 1)])
def test_neumf_and_mlp(pairwise, loss, n_towers):
    """NeuMF.py:69-104 (pairwise: the second _create_inference call makes a second set of dense layers) and, with
    n_towers = 1 on the pairwise graph, MLP.py:45-87 (one tf.layers.Dense stack shared by both items)."""
    rs = np.random.RandomState(1)
    nu, ni, mf_dim, mlp_dim, layers = 8, 10, 5, 4, [8, 6, 3]
    P = {"mf_user": rs.randn(nu, mf_dim), "mf_item": rs.randn(ni, mf_dim), "mlp_user": rs.randn(nu, mlp_dim),
         "mlp_item": rs.randn(ni, mlp_dim)}
    lay, tsz, total = tf_math.ncf_dense_layout(mlp_dim, layers, n_towers)
    P["dense"] = rs.randn(total) * 0.5
    u, i = _ids(rs, 16, nu), _ids(rs, 16, ni)
    third = _ids(rs, 16, ni) if pairwise else (rs.rand(16) < 0.5).astype(np.float64)
    l, G, _, _ = tf_math.ncf_grad(P, u, i, third, pairwise, loss, 0.02, 0.05, mlp_dim, layers, n_towers)
    tp = {k: T(v) for k, v in P.items()}

    def infer(items, tower):
        p, q = tp["mf_user"][I(u)], tp["mf_item"][I(items)]
        m, n = tp["mlp_user"][I(u)], tp["mlp_item"][I(items)]
        h = torch.cat([m, n], 1)
        for (wo, inn, out, bo) in lay:
            W = tp["dense"][tower * tsz + wo:tower * tsz + wo + inn * out].reshape(inn, out)
            b = tp["dense"][tower * tsz + bo:tower * tsz + bo + out]
            h = torch.relu(h @ W + b)
        return p, q, m, n, torch.cat([p * q, h], 1).sum(1)
    p1, q1, m1, n1, out = infer(i, 0)
    if pairwise:
        _, q2, _, n2, out_neg = infer(third, 1 if n_towers == 2 else 0)
        total_t = _pair_loss(loss, out - out_neg) + 0.02 * _l2(p1, q2, q1) + 0.05 * _l2(m1, n2, n1)
    else:
        total_t = _point_loss(loss, torch.as_tensor(third), out) + 0.02 * _l2(p1, q1) + 0.05 * _l2(m1, n1)
    total_t.backward()
    assert abs(float(total_t.detach()) - float(l)) < 1e-9 * max(1.0, abs(float(total_t.detach())))
    for k in P:
        assert np.allclose(G[k], tp[k].grad.numpy(), rtol=1e-9, atol=1e-10), k


def test_lightgcn():
    """LightGCN.py:132-166: mean over layers of A^k E_0, BPR sum, reg on the layer-0 rows."""
    rs = np.random.RandomState(2)
    nu, ni, d, L = 7, 9, 5, 3
    rows = [np.sort(rs.choice(ni, rs.randint(1, 4), replace=False)) for _ in range(nu)]
    ptr = np.cumsum([0] + [len(r) for r in rows]); idx = np.concatenate(rows)
    A = tf_math.lightgcn_adj(ptr, idx, nu, ni, "pre").astype(np.float64)
    e0 = rs.randn(nu + ni, d)
    u, i, j = _ids(rs, 12, nu), _ids(rs, 12, ni), _ids(rs, 12, ni)
    mf, emb, dE0, _ = tf_math.lightgcn_grad(A.astype(np.float32), A.T.tocsr().astype(np.float32), e0.astype(np.float32), nu, u, i, j, 0.01, L)
    te = T(e0)
    tA = torch.tensor(A.toarray(), dtype=torch.float64)
    layers = [te]
    for _ in range(L):
        layers.append(tA @ layers[-1])
    e = torch.stack(layers, 1).mean(1)
    ue, ie = e[:nu], e[nu:]
    x = (ue[I(u)] * ie[I(i)]).sum(1) - (ue[I(u)] * ie[I(j)]).sum(1)
    total = -torch.nn.functional.logsigmoid(x).sum() + 0.01 * _l2(te[I(u)], te[nu + I(i)], te[nu + I(j)])
    total.backward()
    assert abs(float(total.detach()) - (float(mf) + float(emb))) < 1e-5 * abs(float(total.detach()))
    assert np.allclose(dE0, te.grad.numpy(), rtol=1e-4, atol=2e-6)                 # the oracle path is fp32


def test_ngcf():
    """NGCF.py:160-202 + 94-110 with the dropout masks held fixed."""
    rs = np.random.RandomState(3)
    nu, ni, d, sizes = 6, 8, 5, (4, 3)
    rows = [np.sort(rs.choice(ni, rs.randint(1, 4), replace=False)) for _ in range(nu)]
    ptr = np.cumsum([0] + [len(r) for r in rows]); idx = np.concatenate(rows)
    A = tf_math.ngcf_adj(ptr, idx, nu, ni, "norm").astype(np.float64)
    e0 = rs.randn(nu + ni, d)
    W = [[w.astype(np.float64) for w in layer] for layer in tf_math.ngcf_init_weights(rs, d, list(sizes))]
    masks = [(rs.rand(nu + ni, s) < 0.9).astype(np.float64) for s in sizes]
    u, i, j = _ids(rs, 10, nu), _ids(rs, 10, ni), _ids(rs, 10, ni)
    mf, emb, dE0, grads, _ = tf_math.ngcf_loss_and_grad(A, A.T.tocsr(), e0, W, nu, u, i, j, 0.05, masks, keep=0.9)
    te = T(e0)
    tW = [[T(w) for w in layer] for layer in W]
    tA = torch.tensor(A.toarray(), dtype=torch.float64)
    ego, all_e = te, [te]
    for k, (Wgc, bgc, Wbi, bbi) in enumerate(tW):
        side = tA @ ego
        s = torch.nn.functional.leaky_relu(side @ Wgc + bgc, 0.2) + torch.nn.functional.leaky_relu((ego * side) @ Wbi + bbi, 0.2)
        ego = s * torch.as_tensor(masks[k]) / 0.9                                   # tf.nn.dropout(keep_prob)
        all_e.append(torch.nn.functional.normalize(ego, p=2, dim=1, eps=1e-12))
    e = torch.cat(all_e, 1)
    pu, qi, qj = e[I(u)], e[nu + I(i)], e[nu + I(j)]
    total = torch.nn.functional.softplus(-((pu * qi).sum(1) - (pu * qj).sum(1))).sum() + 0.05 * _l2(pu, qi, qj)
    total.backward()
    assert abs(float(total.detach()) - (float(mf) + float(emb))) < 1e-9 * abs(float(total.detach()))
    assert np.allclose(dE0, te.grad.numpy(), rtol=1e-8, atol=1e-10)
    for layer_g, layer_t in zip(grads, tW):
        for g, t in zip(layer_g, layer_t):
            assert np.allclose(g, t.grad.numpy().reshape(g.shape), rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("loss", ["bpr", "hinge", "square"])
def test_sbpr(loss):              # social_recommender/SBPR.py:66-92
    rs = np.random.RandomState(4)
    U, V, B = rs.randn(7, 5), rs.randn(12, 5), rs.randn(12)
    u, i, k, j = _ids(rs, 14, 7), _ids(rs, 14, 12), _ids(rs, 14, 12), _ids(rs, 14, 12)
    s = rs.randint(1, 5, 14).astype(np.float64)
    l, gU, gV, gB, _, _ = tf_math.sbpr_grad(U, V, B, u, i, k, j, s, loss, 0.03)
    tU, tV, tB = T(U), T(V), T(B)
    inf = lambda it: (tU[I(u)], tV[I(it)], tB[I(it)], (tU[I(u)] * tV[I(it)]).sum(1) + tB[I(it)])
    p1, q1, b1, out = inf(i)
    _, q2, b2, out_s = inf(k)
    _, q3, b3, out_n = inf(j)
    total = _pair_loss(loss, (out - out_s) / torch.as_tensor(s)) + _pair_loss(loss, out_s - out_n) + 0.03 * _l2(p1, q2, q1, q3, b1, b2, b3)
    total.backward()
    assert abs(float(total.detach()) - float(l)) < 1e-5 * abs(float(total.detach()))
    for g, t in ((gU, tU), (gV, tV), (gB, tB)):
        assert np.allclose(g, t.grad.numpy(), rtol=2e-5, atol=2e-6)                # sbpr_grad computes in fp32


@pytest.mark.parametrize("act", ["sigmoid", "tanh", "relu", "elu", "identity", "selu"])
def test_spectralcf(act):         # SpectralCF.py:63-91
    rs = np.random.RandomState(5)
    nu, ni, d, K = 6, 9, 4, 2
    A = rs.randn(nu + ni, nu + ni) * 0.3
    e0 = rs.randn(nu + ni, d) * 0.5
    W = [rs.randn(d, d) * 0.5 for _ in range(K)]
    u, i, j = _ids(rs, 10, nu), _ids(rs, 10, ni), _ids(rs, 10, ni)
    l, dE0, dW, _ = tf_math.spectralcf_loss_and_grad(A, e0, W, nu, u, i, j, 0.02, "bpr", act)
    te, tW, tA = T(e0), [T(w) for w in W], torch.tensor(A, dtype=torch.float64)
    f = {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu, "elu": torch.nn.functional.elu,
         "identity": lambda x: x, "selu": torch.selu}[act]
    emb, all_e = te, [te]
    for w in tW:
        emb = f((tA @ emb) @ w)
        all_e.append(emb)
    e = torch.cat(all_e, 1)
    pu, qi, qj = e[I(u)], e[nu + I(i)], e[nu + I(j)]
    total = _pair_loss("bpr", (pu * qi).sum(1) - (pu * qj).sum(1)) + 0.02 * _l2(pu, qi, qj)
    total.backward()
    assert abs(float(total.detach()) - float(l)) < 1e-9 * abs(float(total.detach()))
    assert np.allclose(dE0, te.grad.numpy(), rtol=1e-8, atol=1e-10)
    for g, t in zip(dW, tW):
        assert np.allclose(g, t.grad.numpy(), rtol=1e-8, atol=1e-10)


def test_optimizer_recursions_against_torch():
    """The TF-1.12 update rules of oracle/tf_math.opt_apply against torch.optim where the two libraries define the same
    recursion: momentum (non-Nesterov) exactly; Adam with TF's epsilon placement rewritten for torch (eps_hat vs eps:
    identical m / v recursions and bias correction, checked with eps -> 0); Adagrad with TF's initial accumulator."""
    rs = np.random.RandomState(6)
    x0 = rs.randn(5, 3).astype(np.float32)
    grads = [rs.randn(5, 3).astype(np.float32) for _ in range(6)]

    def run_torch(opt_ctor):
        p = torch.nn.Parameter(torch.tensor(x0.astype(np.float64)))
        opt = opt_ctor([p])
        for g in grads:
            p.grad = torch.tensor(g.astype(np.float64))
            opt.step()
        return p.detach().numpy()

    def run_oracle(kind, hyper_of_step, s0, s1):
        var = x0.copy()
        for t, g in enumerate(grads):
            tf_math.opt_apply(kind, var, g.copy(), s0, s1, None, hyper_of_step(t), dense_var=True)
        return var
    want = run_torch(lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9))
    got = run_oracle("momentum", lambda t: [0.05, 0.9], np.zeros_like(x0), None)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    want = run_torch(lambda ps: torch.optim.Adam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-30))
    lr_t = tf_math.adam_lr_t(1e-2, len(grads))
    got = run_oracle("adam", lambda t: [lr_t[t], 0.9, 0.999, 1e-30], np.zeros_like(x0), np.zeros_like(x0))
    assert np.allclose(got, want, rtol=2e-5, atol=2e-6)
    want = run_torch(lambda ps: torch.optim.Adagrad(ps, lr=0.1, initial_accumulator_value=1e-8, eps=0.0))
    got = run_oracle("adagrad", lambda t: [0.1], np.full_like(x0, 1e-8), None)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
