"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of the reference's TensorFlow-1.12 graphs.

Parity status: UNPINNED at the TensorFlow boundary.  tensorflow==1.12.3 (requirements.txt:4 of
the reference) is a third-party dependency that is neither vendored in /root/reference nor
installable here (no network, Python 3.12), and the reference ships no tests or golden values
for it.  The arithmetic below restates TF's published algorithms; every function cites the
reference call site it stands in for.  What IS checked independently: every hand-derived gradient
agrees with torch.autograd on the same graphs (tests/test_oracle_autograd.py) and with finite
differences (tests/test_oracle.py, tests/test_extras.py); the momentum / Adam / Adagrad recursions
agree with torch.optim where both libraries define the same rule.

Restated call sites (paths relative to /root/reference):
  model/general_recommender/MF.py:54-76        BPRMF / pointwise "GMF" graph
  util/learner.py:2-41                         optimizer / pairwise_loss / pointwise_loss
  util/tool.py:216-224                         l2_loss, log_loss
  model/general_recommender/NeuMF.py:69-104    NeuMF graph (see neumf_*)
  model/general_recommender/MLP.py:45-87       MLP graph
  model/general_recommender/LightGCN.py:35-78,132-166   adjacency + propagation + loss
  model/general_recommender/NGCF.py:94-110,160-202,299-332   NGCF layers, loss, adjacency
  model/general_recommender/APR.py:92-118       adversarial deltas (l2_normalize * eps)
  model/social_recommender/SBPR.py:66-92        SBPR quadruple loss with item bias
  model/general_recommender/SpectralCF.py:37-43,63-91,108-128   spectral operator, layers, loss
TensorFlow pieces (python/training/{adam,adagrad,rmsprop,momentum,gradient_descent}.py,
core/kernels/training_ops.cc, python/ops/nn_impl.py, python/ops/losses/losses_impl.py):
  * embedding_lookup gradients are IndexedSlices; several lookups of one variable are
    concatenated and de-duplicated by summation (optimizer.py::_deduplicate_indexed_slices)
    before the update;
  * Adam on IndexedSlices (adam.py::_apply_sparse_shared) assigns m*b1 and v*b2 over the WHOLE
    variable, scatter-adds the scaled gradient, then updates the WHOLE variable;
    beta powers are fp32 variables multiplied by beta each step (adam.py::_finish);
  * sigmoid_cross_entropy = mean over the batch of max(x,0) - x*z + log1p(exp(-|x|)).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------
def softplus_neg(x):
    """softplus(-x) = -log_sigmoid(x), fp32, overflow-safe (learner.py:22 / tool.py:224)."""
    x = np.asarray(x, dtype=f32)
    pos = x >= 0
    out = np.empty_like(x)
    out[pos] = np.log1p(np.exp(-x[pos]))
    out[~pos] = -x[~pos] + np.log1p(np.exp(x[~pos]))
    return out


def pairwise_loss_and_grad(kind, x):
    """learner.py:18-29 -> (per-sample loss, dloss/dx), fp32."""
    x = np.asarray(x, dtype=f32)
    if kind == "bpr":
        return softplus_neg(x), (f32(-1.0) / (f32(1.0) + np.exp(x))).astype(f32)
    if kind == "hinge":  # sum(max(y + margin, 0)) exactly as written in the reference
        t = x + f32(1.0)
        return np.maximum(t, f32(0)), (t > 0).astype(f32)
    if kind == "square":
        t = f32(1.0) - x
        return t * t, f32(-2.0) * t
    raise Exception("please choose a suitable loss function")


def pointwise_loss_and_grad(kind, z, x):
    """learner.py:31-41 -> (per-sample loss contribution, dloss/dx); CE already / batch."""
    x = np.asarray(x, dtype=f32)
    z = np.asarray(z, dtype=f32)
    if kind == "cross_entropy":
        inv_b = f32(1.0) / f32(len(x))
        e = np.exp(-np.abs(x))
        l = (np.maximum(x, f32(0)) - x * z + np.log1p(e)) * inv_b
        s = np.where(x >= 0, f32(1.0) / (f32(1.0) + e), e / (f32(1.0) + e)).astype(f32)
        return l.astype(f32), ((s - z) * inv_b).astype(f32)
    if kind == "square":
        t = z - x
        return t * t, f32(-2.0) * t
    raise Exception("please choose a suitable loss function")


# ----------------------------------------------------------------------------------------
# MF graph (MF.py:54-72)
# ----------------------------------------------------------------------------------------
def mf_pairwise_grad(U, V, users, pos, neg, loss="bpr", reg=0.0):
    """-> (loss, gU, gV, touchedU, touchedV).  Gradients are the de-duplicated (summed)
    IndexedSlices scattered into dense fp32 arrays."""
    U = np.asarray(U, f32); V = np.asarray(V, f32)
    pu, qi, qj = U[users], V[pos], V[neg]
    x = (pu * qi).sum(1, dtype=f32) - (pu * qj).sum(1, dtype=f32)      # MF.py:59,66
    l, g = pairwise_loss_and_grad(loss, x)
    reg = f32(reg)
    total = l.sum(dtype=f32) + reg * f32(0.5) * ((pu * pu).sum(dtype=f32) + (qj * qj).sum(dtype=f32)
                                                 + (qi * qi).sum(dtype=f32))  # MF.py:67
    gU = np.zeros_like(U); gV = np.zeros_like(V)
    g = g[:, None]
    np.add.at(gU, users, (g * (qi - qj) + reg * pu).astype(f32))
    np.add.at(gV, pos, (g * pu + reg * qi).astype(f32))
    np.add.at(gV, neg, (-g * pu + reg * qj).astype(f32))
    tU = np.zeros(U.shape[0], bool); tU[users] = True
    tV = np.zeros(V.shape[0], bool); tV[pos] = True; tV[neg] = True
    return f32(total), gU, gV, tU, tV


def mf_pointwise_grad(U, V, users, items, labels, loss="cross_entropy", reg=0.0):
    U = np.asarray(U, f32); V = np.asarray(V, f32)
    pu, qi = U[users], V[items]
    x = (pu * qi).sum(1, dtype=f32)
    l, g = pointwise_loss_and_grad(loss, labels, x)
    reg = f32(reg)
    total = l.sum(dtype=f32) + reg * f32(0.5) * ((pu * pu).sum(dtype=f32) + (qi * qi).sum(dtype=f32))
    gU = np.zeros_like(U); gV = np.zeros_like(V)
    g = g[:, None]
    np.add.at(gU, users, (g * qi + reg * pu).astype(f32))
    np.add.at(gV, items, (g * pu + reg * qi).astype(f32))
    tU = np.zeros(U.shape[0], bool); tU[users] = True
    tV = np.zeros(V.shape[0], bool); tV[items] = True
    return f32(total), gU, gV, tU, tV


# ----------------------------------------------------------------------------------------
# optimizers (learner.py:2-15)
# ----------------------------------------------------------------------------------------
def adam_lr_t(lr, steps, beta1=0.9, beta2=0.999, start_step=0):
    """Per-step lr_t = lr*sqrt(1-b2^t)/(1-b1^t) with fp32 beta-power variables that are
    multiplied by beta after every step (adam.py::_finish / _prepare)."""
    b1, b2, lr = f32(beta1), f32(beta2), f32(lr)
    p1, p2 = b1, b2
    out = np.empty(start_step + steps, dtype=f32)
    for t in range(start_step + steps):
        out[t] = lr * np.sqrt(f32(1.0) - p2) / (f32(1.0) - p1)
        p1 = f32(p1 * b1)
        p2 = f32(p2 * b2)
    return out[start_step:]


DEFAULT_HYPER = {
    # learner.py:4-14 defaults of the TF-1.12 constructors
    "gd": lambda lr: [lr],
    "adam": lambda lr: [lr, 0.9, 0.999, 1e-8],      # h[0] is replaced by lr_t per step
    "adagrad": lambda lr: [lr],                     # initial_accumulator_value=1e-8
    "rmsprop": lambda lr: [lr, 0.9, 0.0, 1e-10],    # decay, momentum, epsilon
    "momentum": lambda lr: [lr, 0.9],
}
SLOT_INIT = {"gd": (None, None), "adam": (0.0, 0.0), "adagrad": (1e-8, None),
             "rmsprop": (1.0, 0.0), "momentum": (0.0, None)}  # rmsprop ms starts at ones


def opt_apply(kind, var, g, s0, s1, touched, hyper, dense_var=False):
    """In-place TF-1.12 update of one variable; mirrors csrc/optim.cu operation by operation.
    `touched` is a bool row mask (ignored for dense variables / gd / adam)."""
    h = [f32(v) for v in list(hyper) + [0.0] * (4 - len(hyper))]
    one = f32(1.0)
    if kind == "gd":
        var -= g * h[0]
    elif kind == "adam":
        omb1, omb2 = one - h[1], one - h[2]
        if dense_var:
            s0 += (g - s0) * omb1
            s1 += (g * g - s1) * omb2
            var -= (s0 * h[0]) / (np.sqrt(s1) + h[3])
        else:
            s0[...] = s0 * h[1] + g * omb1
            s1[...] = s1 * h[2] + (g * g) * omb2
            var -= (h[0] * s0) / (np.sqrt(s1) + h[3])
    else:
        rows = slice(None) if (dense_var or touched is None) else touched
        gg = g[rows]
        if kind == "adagrad":
            s0[rows] = s0[rows] + gg * gg
            var[rows] = var[rows] - (h[0] * gg) * (one / np.sqrt(s0[rows]))
        elif kind == "rmsprop":
            if dense_var:     # ApplyRMSProp functor (training_ops.cc)
                s0[rows] = s0[rows] + (gg * gg - s0[rows]) * (one - h[1])
                s1[rows] = s1[rows] * h[2] + (h[0] * gg) * (one / np.sqrt(s0[rows] + h[3]))
            else:             # SparseApplyRMSProp: ms*rho + g*g*(1-rho); mom*mu + rsqrt(ms+eps)*lr*g
                s0[rows] = s0[rows] * h[1] + (gg * gg) * (one - h[1])
                s1[rows] = s1[rows] * h[2] + ((one / np.sqrt(s0[rows] + h[3])) * h[0]) * gg
            var[rows] = var[rows] - s1[rows]
        elif kind == "momentum":
            s0[rows] = s0[rows] * h[1] + gg
            var[rows] = var[rows] - s0[rows] * h[0]
        else:
            raise ValueError("please select a suitable optimizer")
    return var


class MFTrainer:
    """CPU stand-in for MF.build_graph + the sess.run((loss, optimizer)) loop (MF.py:78-108)."""

    def __init__(self, U, V, learner="adam", lr=1e-3, loss="bpr", reg=0.0, pairwise=True):
        self.U = np.array(U, dtype=f32); self.V = np.array(V, dtype=f32)
        self.learner, self.lr, self.loss, self.reg, self.pairwise = learner, lr, loss, reg, pairwise
        i0, i1 = SLOT_INIT[learner]
        mk = lambda a, v: None if v is None else np.full_like(a, v)
        self.s0U, self.s1U = mk(self.U, i0), mk(self.U, i1)
        self.s0V, self.s1V = mk(self.V, i0), mk(self.V, i1)
        self.t = 0

    def step(self, users, items, third):
        if self.pairwise:
            l, gU, gV, tU, tV = mf_pairwise_grad(self.U, self.V, users, items, third, self.loss, self.reg)
        else:
            l, gU, gV, tU, tV = mf_pointwise_grad(self.U, self.V, users, items, third, self.loss, self.reg)
        hyper = DEFAULT_HYPER[self.learner](self.lr)
        if self.learner == "adam":
            hyper[0] = adam_lr_t(self.lr, 1, start_step=self.t)[0]
        opt_apply(self.learner, self.U, gU, self.s0U, self.s1U, tU, hyper)
        opt_apply(self.learner, self.V, gV, self.s0V, self.s1V, tV, hyper)
        self.t += 1
        return l

    def epoch(self, users, items, third, batch_size):
        n = len(users)
        losses = []
        for off in range(0, n, batch_size):
            sl = slice(off, min(n, off + batch_size))
            losses.append(self.step(users[sl], items[sl], third[sl]))
        return np.asarray(losses, dtype=f32)


# ----------------------------------------------------------------------------------------
# APR (APR.py:92-118) and SBPR (social_recommender/SBPR.py:66-92) -- SURVEY.md 8(f) rank 3
# ----------------------------------------------------------------------------------------
def l2_normalize_rows(x, scale):
    """tf.nn.l2_normalize(x, 1) * scale: x * rsqrt(max(sum(x^2), 1e-12)) * scale (APR.py:103-104,117-118)."""
    x = np.asarray(x, f32)
    ss = (x * x).sum(1, dtype=f32, keepdims=True)
    return ((x * (f32(1.0) / np.sqrt(np.maximum(ss, f32(1e-12))))) * f32(scale)).astype(f32)


def sbpr_grad(U, V, B, users, pos, soc, neg, suk, loss="bpr", reg=0.0):
    """SBPR._create_loss (SBPR.py:79-92) -> (loss, gU, gV, gB, touchedU, touchedV); x = <p, q> + b;
    loss = l((x_i - x_k) / s) + l(x_k - x_j) + reg * l2_loss(p1, q2, q1, q3, b1, b2, b3)."""
    U = np.asarray(U, f32); V = np.asarray(V, f32); B = np.asarray(B, f32)
    s = np.asarray(suk, f32)
    pu, qi, qk, qj = U[users], V[pos], V[soc], V[neg]
    bi, bk, bj = B[pos], B[soc], B[neg]
    xi = (pu * qi).sum(1, dtype=f32) + bi
    xk = (pu * qk).sum(1, dtype=f32) + bk
    xj = (pu * qj).sum(1, dtype=f32) + bj
    l1, g1 = pairwise_loss_and_grad(loss, (xi - xk) / s)
    l2, g2 = pairwise_loss_and_grad(loss, xk - xj)
    reg = f32(reg)
    sq = sum((a * a).sum(dtype=f32) for a in (pu, qk, qi, qj, bi, bk, bj))
    total = l1.sum(dtype=f32) + l2.sum(dtype=f32) + reg * f32(0.5) * f32(sq)
    ci = (g1 / s).astype(f32); ck = (g2 - ci).astype(f32); cj = (-g2).astype(f32)
    gU = np.zeros_like(U); gV = np.zeros_like(V); gB = np.zeros_like(B)
    np.add.at(gU, users, (ci[:, None] * qi + ck[:, None] * qk + cj[:, None] * qj + reg * pu).astype(f32))
    np.add.at(gV, pos, (ci[:, None] * pu + reg * qi).astype(f32))
    np.add.at(gV, soc, (ck[:, None] * pu + reg * qk).astype(f32))
    np.add.at(gV, neg, (cj[:, None] * pu + reg * qj).astype(f32))
    np.add.at(gB, pos, (ci + reg * bi).astype(f32))
    np.add.at(gB, soc, (ck + reg * bk).astype(f32))
    np.add.at(gB, neg, (cj + reg * bj).astype(f32))
    tU = np.zeros(U.shape[0], bool); tU[users] = True
    tV = np.zeros(V.shape[0], bool); tV[pos] = True; tV[soc] = True; tV[neg] = True
    return f32(total), gU, gV, gB, tU, tV


class SBPRTrainer:
    """CPU stand-in for SBPR.build_graph + the sess.run((loss, optimizer)) loop (SBPR.py:94-121)."""

    def __init__(self, U, V, B, learner="adam", lr=1e-3, loss="bpr", reg=0.01):
        self.U = np.array(U, dtype=f32); self.V = np.array(V, dtype=f32); self.B = np.array(B, dtype=f32)
        self.learner, self.lr, self.loss, self.reg = learner, lr, loss, reg
        i0, i1 = SLOT_INIT[learner]
        mk = lambda a, v: None if v is None else np.full_like(a, v)
        self.slots = [(mk(a, i0), mk(a, i1)) for a in (self.U, self.V, self.B)]
        self.t = 0

    def step(self, users, pos, soc, neg, suk):
        l, gU, gV, gB, tU, tV = sbpr_grad(self.U, self.V, self.B, users, pos, soc, neg, suk, self.loss, self.reg)
        hyper = DEFAULT_HYPER[self.learner](self.lr)
        if self.learner == "adam":
            hyper[0] = adam_lr_t(self.lr, 1, start_step=self.t)[0]
        for var, g, (s0, s1), touched in ((self.U, gU, self.slots[0], tU), (self.V, gV, self.slots[1], tV),
                                          (self.B, gB, self.slots[2], tV)):
            opt_apply(self.learner, var, g, s0, s1, touched, hyper)
        self.t += 1
        return l

    def epoch(self, users, pos, soc, neg, suk, batch_size):
        n = len(users)
        losses = []
        for off in range(0, n, batch_size):
            sl = slice(off, min(n, off + batch_size))
            losses.append(self.step(users[sl], pos[sl], soc[sl], neg[sl], suk[sl]))
        return np.asarray(losses, dtype=f32)


# ----------------------------------------------------------------------------------------
# SpectralCF (general_recommender/SpectralCF.py:37-43, 63-91, 108-128) -- SURVEY.md 8(f) rank 3
# ----------------------------------------------------------------------------------------
def spectralcf_a_hat(train_indptr, train_indices, num_users, num_items):
    """The dense spectral operator of SpectralCF: A = I + bipartite adjacency (:108-114), D = row sums (:116-119),
    L = I - D^-1 A (:121-128), (lamda, U) = eig(L) (:41-42), A_hat = U U^T + U diag(lamda) U^T cast to fp32 (:67-69;
    a complex result of eig loses its imaginary part in that cast, as in the reference)."""
    nu, ni = int(num_users), int(num_items)
    n = nu + ni
    graph = np.zeros((nu, ni), dtype=f32)
    rows = np.repeat(np.arange(nu), np.diff(np.asarray(train_indptr)))
    graph[rows, np.asarray(train_indices)] = 1.0
    A = np.zeros((n, n), dtype=f32)
    A[:nu, nu:] = graph
    A[nu:, :nu] = graph.T
    A = np.identity(n, dtype=f32) + A
    D = A.sum(axis=1)
    L = np.identity(n, dtype=f32) - np.dot(np.diag(np.power(D, -1)), A)
    lamda, U = np.linalg.eig(L)
    lamda = np.diag(lamda)
    A_hat = np.dot(U, U.T) + np.dot(np.dot(U, lamda), U.T)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.ascontiguousarray(A_hat.astype(f32))


_SELU_SCALE, _SELU_ALPHA = 1.0507009873554805, 1.6732632423543772


def activation(act, z):
    """tool.activation_function (util/tool.py:10-33) -> fp value of the same dtype."""
    z = np.asarray(z)
    one = z.dtype.type(1)
    if act == "sigmoid":
        return one / (one + np.exp(-z))
    if act == "tanh":
        return np.tanh(z)
    if act == "relu":
        return np.maximum(z, 0)
    if act == "elu":
        return np.where(z > 0, z, np.exp(np.minimum(z, 0)) - one)
    if act == "identity":
        return z
    if act == "selu":
        return (z.dtype.type(_SELU_SCALE) * np.where(z > 0, z, z.dtype.type(_SELU_ALPHA) * (np.exp(np.minimum(z, 0)) - one))).astype(z.dtype)
    raise NotImplementedError("ERROR")


def activation_grad_from_output(act, y):
    """d act / d z written in terms of the OUTPUT y (what the backward kernel has at hand)."""
    y = np.asarray(y)
    one = y.dtype.type(1)
    if act == "sigmoid":
        return y * (one - y)
    if act == "tanh":
        return one - y * y
    if act == "relu":
        return (y > 0).astype(y.dtype)
    if act == "elu":
        return np.where(y > 0, one, y + one)
    if act == "identity":
        return np.ones_like(y)
    if act == "selu":
        return np.where(y > 0, y.dtype.type(_SELU_SCALE), y + y.dtype.type(_SELU_SCALE * _SELU_ALPHA))
    raise NotImplementedError("ERROR")


def spectralcf_forward(A_hat, e0, filters, act="sigmoid"):
    """SpectralCF._create_inference (:63-83): E_k = act((A_hat E_{k-1}) W_k); returns (all_emb [N, d (K+1)], [S_k])."""
    emb = np.asarray(e0)
    all_emb, sides = [emb], []
    for W in filters:
        side = A_hat.astype(emb.dtype) @ emb
        emb = activation(act, side @ W)
        sides.append(side)
        all_emb.append(emb)
    return np.concatenate(all_emb, axis=1), sides


def spectralcf_loss_and_grad(A_hat, e0, filters, num_users, users, pos, neg, reg, loss="bpr", act="sigmoid"):
    """Loss (:85-91: pairwise_loss on the concatenated rows + reg * l2_loss(u, i, j)) and its gradient w.r.t. the
    embedding table and every filter -- all dense (they flow through tf.matmul with A_hat)."""
    A = A_hat.astype(np.asarray(e0).dtype)
    all_emb, sides = spectralcf_forward(A, e0, filters, act)
    dt = all_emb.dtype.type
    d = e0.shape[1]
    ue, ie = all_emb[:num_users], all_emb[num_users:]
    pu, qi, qj = ue[users], ie[pos], ie[neg]
    x = (pu * qi).sum(1) - (pu * qj).sum(1)
    if all_emb.dtype == f32:
        l, g = pairwise_loss_and_grad(loss.lower(), x)
    else:
        l = {"bpr": np.log1p(np.exp(-x)), "hinge": np.maximum(x + 1, 0), "square": (1 - x) ** 2}[loss.lower()]
        g = {"bpr": -1 / (1 + np.exp(x)), "hinge": (x + 1 > 0).astype(x.dtype), "square": -2 * (1 - x)}[loss.lower()]
    reg = dt(reg)
    total = l.sum() + reg * dt(0.5) * ((pu * pu).sum() + (qi * qi).sum() + (qj * qj).sum())
    G = np.zeros_like(all_emb)
    g = g[:, None].astype(all_emb.dtype)
    np.add.at(G, users, g * (qi - qj) + reg * pu)
    np.add.at(G, num_users + np.asarray(pos), g * pu + reg * qi)
    np.add.at(G, num_users + np.asarray(neg), -g * pu + reg * qj)
    K = len(filters)
    dW = [None] * K
    carry = np.zeros((all_emb.shape[0], d), dtype=all_emb.dtype)
    for k in range(K, 0, -1):
        Ek = all_emb[:, k * d:(k + 1) * d]
        dZ = (G[:, k * d:(k + 1) * d] + carry) * activation_grad_from_output(act, Ek)
        dW[k - 1] = sides[k - 1].T @ dZ
        carry = A.T @ (dZ @ filters[k - 1].T)
    dE0 = G[:, :d] + carry
    return total, dE0, dW, all_emb


class SpectralCFTrainer:
    """SpectralCF.train_model's step on the restatement: dense TF-1.12 optimizer over the table and every filter."""

    def __init__(self, A_hat, e0, filters, num_users, learner="adam", lr=1e-3, reg=1e-3, loss="bpr", act="sigmoid"):
        self.A = np.asarray(A_hat, f32)
        self.e0 = np.array(e0, f32); self.filters = [np.array(W, f32) for W in filters]
        self.nu, self.learner, self.lr, self.reg, self.loss, self.act = num_users, learner, lr, reg, loss, act
        i0, i1 = SLOT_INIT[learner]
        mk = lambda a, v: None if v is None else np.full_like(a, v)
        self.slots = [(mk(a, i0), mk(a, i1)) for a in [self.e0] + self.filters]
        self.t = 0

    def step(self, users, pos, neg):
        l, dE0, dW, _ = spectralcf_loss_and_grad(self.A, self.e0, self.filters, self.nu, users, pos, neg, self.reg,
                                                 self.loss, self.act)
        hyper = DEFAULT_HYPER[self.learner](self.lr)
        if self.learner == "adam":
            hyper[0] = adam_lr_t(self.lr, 1, start_step=self.t)[0]
        for var, g, (s0, s1) in zip([self.e0] + self.filters, [dE0] + dW, self.slots):
            opt_apply(self.learner, var, np.asarray(g, f32), s0, s1, None, hyper, dense_var=True)
        self.t += 1
        return f32(l)

    def embeddings(self):
        return spectralcf_forward(self.A, self.e0, self.filters, self.act)[0]


# ----------------------------------------------------------------------------------------
# LightGCN: LightGCN.py:35-78 (adjacency), 132-149 (propagation), 156-166 (loss)
# ----------------------------------------------------------------------------------------
def lightgcn_adj(train_indptr, train_indices, num_users, num_items, adj_type="pre"):
    """create_adj_mat (LightGCN.py:35-78) with the reference's own scipy calls -> fp32 CSR."""
    import scipy.sparse as sp
    n = num_users + num_items
    users = np.repeat(np.arange(num_users), np.diff(train_indptr)).astype(np.int32)
    items = np.asarray(train_indices, dtype=np.int32)
    tmp = sp.csr_matrix((np.ones(len(users), np.float32), (users, items + num_users)), shape=(n, n))
    adj = tmp + tmp.T

    def single(a):
        rowsum = np.array(a.sum(1))
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.
        return sp.diags(d_inv).dot(a).tocoo()

    if adj_type == "plain":
        m = adj
    elif adj_type == "norm":
        m = single(adj + sp.eye(n))
    elif adj_type == "gcmc":
        m = single(adj)
    elif adj_type == "pre":
        rowsum = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, -0.5).flatten()
        d_inv[np.isinf(d_inv)] = 0.
        d = sp.diags(d_inv)
        m = d.dot(adj).dot(d)
    else:
        m = single(adj) + sp.eye(n)
    m = m.tocoo().astype(np.float32).tocsr()   # _convert_sp_mat_to_sp_tensor (LightGCN.py:151-154)
    m.sort_indices()
    return m


def lightgcn_propagate(A, e0, n_layers):
    """mean over [E_0, A E_0, ...] (LightGCN.py:132-149); scipy's csr @ dense accumulates each
    output row sequentially over its nnz with separate multiply and add, like TF's CPU kernel."""
    e0 = np.asarray(e0, f32)
    layers = [e0]
    x = e0
    for _ in range(n_layers):
        x = (A @ x).astype(f32)
        layers.append(x)
    s = layers[0].copy()
    for y in layers[1:]:
        s = (s + y).astype(f32)
    return (s / f32(n_layers + 1)).astype(f32), layers


def lightgcn_grad(A, AT, e0, num_users, users, pos, neg, reg, n_layers):
    """-> (mf_loss, emb_loss, dE0, e_final): manual backprop of LightGCN.py:99-130."""
    e, _ = lightgcn_propagate(A, e0, n_layers)
    ru, ri, rj = users, num_users + pos, num_users + neg
    x = (e[ru] * e[ri]).sum(1, dtype=f32) - (e[ru] * e[rj]).sum(1, dtype=f32)
    l, g = pairwise_loss_and_grad("bpr", x)
    reg = f32(reg)
    emb = reg * f32(0.5) * ((e0[ru] ** 2).sum(dtype=f32) + (e0[ri] ** 2).sum(dtype=f32) + (e0[rj] ** 2).sum(dtype=f32))
    scale = f32(1.0) / f32(n_layers + 1)
    G = np.zeros_like(e0)
    gs = (g * scale)[:, None].astype(f32)
    np.add.at(G, ru, gs * (e[ri] - e[rj]))
    np.add.at(G, ri, gs * e[ru])
    np.add.at(G, rj, -gs * e[ru])
    R = np.zeros_like(e0)
    for r in (ru, ri, rj):
        np.add.at(R, r, reg * e0[r])
    t = G
    for _ in range(n_layers):
        t = (G + (AT @ t).astype(f32)).astype(f32)
    return l.sum(dtype=f32), f32(emb), (R + t).astype(f32), e


class LightGCNTrainer:
    def __init__(self, A, e0, num_users, n_layers, lr=0.01, reg=1e-3):
        self.A, self.AT = A, A.T.tocsr()
        self.AT.sort_indices()
        self.e0 = np.array(e0, f32)
        self.m, self.v = np.zeros_like(self.e0), np.zeros_like(self.e0)
        self.num_users, self.n_layers, self.lr, self.reg, self.t = num_users, n_layers, lr, reg, 0

    def step(self, users, pos, neg):
        mf, emb, g, _ = lightgcn_grad(self.A, self.AT, self.e0, self.num_users, users, pos, neg, self.reg,
                                      self.n_layers)
        hyper = [adam_lr_t(self.lr, 1, start_step=self.t)[0], 0.9, 0.999, 1e-8]
        opt_apply("adam", self.e0, g, self.m, self.v, None, hyper, dense_var=True)
        self.t += 1
        return mf, emb

    def epoch(self, users, pos, neg, batch_size):
        out = []
        for off in range(0, len(users), batch_size):
            sl = slice(off, min(len(users), off + batch_size))
            out.append(self.step(users[sl], pos[sl], neg[sl]))
        return np.asarray(out, f32)


# ----------------------------------------------------------------------------------------
# NCF family: NeuMF.py:69-104, MLP.py:45-87
# ----------------------------------------------------------------------------------------
def ncf_dense_layout(mlp_dim, layers, n_towers):
    """Packed layout used by the product: per tower, per layer kernel [in,out] then bias."""
    lay, off, inn = [], 0, 2 * mlp_dim
    for out in layers:
        lay.append((off, inn, out, off + inn * out))
        off += inn * out + out
        inn = out
    return lay, off, off * n_towers


def ncf_init_dense(mlp_dim, layers, n_towers, rs):
    """tf.layers.dense defaults: glorot_uniform kernel, zero bias."""
    lay, tower, total = ncf_dense_layout(mlp_dim, layers, n_towers)
    buf = np.zeros(total, f32)
    for t in range(n_towers):
        for (wo, inn, out, bo) in lay:
            lim = np.sqrt(6.0 / (inn + out))
            buf[t * tower + wo:t * tower + wo + inn * out] = rs.uniform(-lim, lim, inn * out).astype(f32)
    return buf


def _ncf_tower(dense, lay, tower_off, h):
    acts = [h]
    for (wo, inn, out, bo) in lay:
        W = dense[tower_off + wo:tower_off + wo + inn * out].reshape(inn, out)
        b = dense[tower_off + bo:tower_off + bo + out]
        h = np.maximum(h @ W + b, 0).astype(h.dtype)
        acts.append(h)
    return acts


def ncf_predict(P, users, items, mlp_dim, layers, tower=0):
    """prediction = reduce_sum(concat(mf_vector, mlp_vector), 1)  (NeuMF.py:85)."""
    lay, tsz, _ = ncf_dense_layout(mlp_dim, layers, 1)
    dt = P["dense"].dtype if layers else P["mf_user"].dtype
    y = np.zeros(len(users), dt)
    if P["mf_user"] is not None and P["mf_user"].shape[1] > 0:
        y = y + (P["mf_user"][users] * P["mf_item"][items]).sum(1)
    if layers:
        h = np.concatenate([P["mlp_user"][users], P["mlp_item"][items]], axis=1)
        acts = _ncf_tower(P["dense"], lay, tower * tsz, h)
        y = y + acts[-1].sum(1)
    return y.astype(dt)


def ncf_grad(P, users, items, third, pairwise, loss, reg_mf, reg_mlp, mlp_dim, layers, n_towers):
    """-> (loss, grads dict with the same keys as P, touched_user, touched_item).
    Manual backprop of the TF graph; dtype follows P (fp32 in parity tests)."""
    lay, tsz, total = ncf_dense_layout(mlp_dim, layers, n_towers)
    has_mf = P["mf_user"] is not None and P["mf_user"].shape[1] > 0
    dt = (P["dense"] if layers else P["mf_user"]).dtype
    G = {k: (np.zeros_like(v) if v is not None else None) for k, v in P.items()}
    reg_mf, reg_mlp = dt.type(reg_mf), dt.type(reg_mlp)
    passes = [(items, 0)] + ([(third, 1 if n_towers == 2 else 0)] if pairwise else [])
    cache, yh = [], []
    for it, tw in passes:
        y = np.zeros(len(users), dt)
        acts = None
        if has_mf:
            y = y + (P["mf_user"][users] * P["mf_item"][it]).sum(1)
        if layers:
            h = np.concatenate([P["mlp_user"][users], P["mlp_item"][it]], axis=1)
            acts = _ncf_tower(P["dense"], lay, tw * tsz, h)
            y = y + acts[-1].sum(1)
        cache.append(acts); yh.append(y.astype(dt))
    if dt == np.float32:
        if pairwise:
            l, g = pairwise_loss_and_grad(loss, yh[0] - yh[1])
        else:
            l, g = pointwise_loss_and_grad(loss, third, yh[0])
    else:  # float64 path for finite-difference pins
        x = yh[0] - yh[1] if pairwise else yh[0]
        if pairwise:
            l = {"bpr": np.log1p(np.exp(-x)), "hinge": np.maximum(x + 1, 0), "square": (1 - x) ** 2}[loss]
            g = {"bpr": -1 / (1 + np.exp(x)), "hinge": (x + 1 > 0) * 1.0, "square": -2 * (1 - x)}[loss]
        elif loss == "cross_entropy":
            l = (np.maximum(x, 0) - x * third + np.log1p(np.exp(-np.abs(x)))) / len(x)
            g = (1 / (1 + np.exp(-x)) - third) / len(x)
        else:
            l = (third - x) ** 2; g = -2 * (third - x)
    gs = [g, -g] if pairwise else [g]
    total_loss = l.sum()
    for p, (it, tw) in enumerate(passes):
        gp = gs[p][:, None].astype(dt)
        if has_mf:
            pu, qi = P["mf_user"][users], P["mf_item"][it]
            np.add.at(G["mf_user"], users, (gp * qi + (reg_mf * pu if p == 0 else 0)).astype(dt))
            np.add.at(G["mf_item"], it, (gp * pu + reg_mf * qi).astype(dt))
            total_loss += reg_mf * dt.type(0.5) * ((qi * qi).sum() + ((pu * pu).sum() if p == 0 else 0))
        if layers:
            acts = cache[p]
            delta = (gp * (acts[-1] > 0)).astype(dt)
            for li in range(len(lay) - 1, -1, -1):
                wo, inn, out, bo = lay[li]
                W = P["dense"][tw * tsz + wo:tw * tsz + wo + inn * out].reshape(inn, out)
                G["dense"][tw * tsz + wo:tw * tsz + wo + inn * out] += (acts[li].T @ delta).reshape(-1).astype(dt)
                G["dense"][tw * tsz + bo:tw * tsz + bo + out] += delta.sum(0)
                delta = (delta @ W.T).astype(dt)
                if li > 0:
                    delta = (delta * (acts[li] > 0)).astype(dt)
            mu, mi = P["mlp_user"][users], P["mlp_item"][it]
            np.add.at(G["mlp_user"], users, (delta[:, :mlp_dim] + (reg_mlp * mu if p == 0 else 0)).astype(dt))
            np.add.at(G["mlp_item"], it, (delta[:, mlp_dim:] + reg_mlp * mi).astype(dt))
            total_loss += reg_mlp * dt.type(0.5) * ((mi * mi).sum() + ((mu * mu).sum() if p == 0 else 0))
    nu = (P["mlp_user"] if layers else P["mf_user"]).shape[0]
    ni = (P["mlp_item"] if layers else P["mf_item"]).shape[0]
    tU = np.zeros(nu, bool); tU[users] = True
    tI = np.zeros(ni, bool)
    for it, _ in passes:
        tI[it] = True
    return dt.type(total_loss), G, tU, tI


class NCFTrainer:
    """CPU stand-in for NeuMF/MLP build_graph + train loop (NeuMF.py:106-151)."""
    TABLES = ("mf_user", "mf_item", "mlp_user", "mlp_item")

    def __init__(self, P, mlp_dim, layers, n_towers, learner="adam", lr=1e-3, loss="cross_entropy",
                 reg_mf=0.0, reg_mlp=0.0, pairwise=False):
        self.P = {k: (np.array(v, dtype=f32) if v is not None else None) for k, v in P.items()}
        self.mlp_dim, self.layers, self.n_towers = mlp_dim, list(layers), n_towers
        self.learner, self.lr, self.loss, self.pairwise = learner, lr, loss, pairwise
        self.reg_mf, self.reg_mlp = reg_mf, reg_mlp
        i0, i1 = SLOT_INIT[learner]
        mk = lambda a, v: None if (v is None or a is None) else np.full_like(a, v)
        self.s0 = {k: mk(v, i0) for k, v in self.P.items()}
        self.s1 = {k: mk(v, i1) for k, v in self.P.items()}
        self.t = 0

    def step(self, users, items, third):
        l, G, tU, tI = ncf_grad(self.P, users, items, third, self.pairwise, self.loss, self.reg_mf,
                                self.reg_mlp, self.mlp_dim, self.layers, self.n_towers)
        hyper = DEFAULT_HYPER[self.learner](self.lr)
        if self.learner == "adam":
            hyper[0] = adam_lr_t(self.lr, 1, start_step=self.t)[0]
        for k in self.TABLES:
            if self.P[k] is None or self.P[k].shape[1] == 0:
                continue
            opt_apply(self.learner, self.P[k], G[k], self.s0[k], self.s1[k], tU if "user" in k else tI, hyper)
        if self.layers:
            opt_apply(self.learner, self.P["dense"], G["dense"], self.s0["dense"], self.s1["dense"],
                      None, hyper, dense_var=True)
        self.t += 1
        return l

    def epoch(self, users, items, third, batch_size):
        out = []
        for off in range(0, len(users), batch_size):
            sl = slice(off, min(len(users), off + batch_size))
            out.append(self.step(users[sl], items[sl], third[sl]))
        return np.asarray(out, f32)


# ----------------------------------------------------------------------------------------
# NGCF: restatement of NGCF.py:160-202 (propagation + dense part) and
# NGCF.py:94-110 (loss), ahead of the CUDA kernels.  Same conventions as above: numpy, the
# arrays' own dtype (fp32 in use, fp64 for the finite-difference checks), explicit dropout masks
# so that a GPU run and this restatement can share them (tf.nn.dropout is ALWAYS on in the
# reference, also at evaluation time, NGCF.py:193).  Parity unpinned at the TensorFlow boundary.
# ----------------------------------------------------------------------------------------
LEAKY_ALPHA = 0.2            # tf.nn.leaky_relu default
L2NORM_EPS = 1e-12           # tf.nn.l2_normalize default epsilon


def ngcf_adj(train_indptr, train_indices, num_users, num_items, adj_type="norm"):
    """get_adj_mat (NGCF.py:299-322): 'norm' = D^-1 (A + I) (the conf default), 'plain', 'gcmc',
    otherwise D^-1 A + I -- the same four matrices create_adj_mat of LightGCN builds."""
    return lightgcn_adj(train_indptr, train_indices, num_users, num_items, adj_type)


def ngcf_init_weights(rs, emb_dim, layer_size):
    """[(W_gc, b_gc, W_bi, b_bi)] per layer with xavier-normal kernels (conf default) and the
    reference's shapes: W [d_k, d_{k+1}], b [1, d_{k+1}] (NGCF.py:262-284)."""
    dims = [emb_dim] + list(layer_size)
    out = []
    for k in range(len(layer_size)):
        std = np.sqrt(2.0 / (dims[k] + dims[k + 1]))
        mk = lambda r, c: (rs.randn(r, c) * std).astype(f32)
        out.append((mk(dims[k], dims[k + 1]), mk(1, dims[k + 1]), mk(dims[k], dims[k + 1]), mk(1, dims[k + 1])))
    return out


def _leaky(x):
    return np.where(x > 0, x, x * x.dtype.type(LEAKY_ALPHA))


def ngcf_forward(A, e0, weights, masks=None, keep=1.0):
    """_create_ngcf_embed (NGCF.py:160-202).  masks[k]: 0/1 array like layer k's output (None = no
    dropout); tf.nn.dropout(x, keep) = x * mask / keep.  The n_fold row slabs (174-179) are a
    memory work-around: concatenating slab products equals one product with the whole matrix.
    -> (all_embeddings [N, d0 + sum(d_k)], cache for ngcf_backward)."""
    dt = e0.dtype.type
    ego = e0
    outs, cache = [e0], []
    for k, (Wgc, bgc, Wbi, bbi) in enumerate(weights):
        side = np.asarray(A @ ego, dtype=e0.dtype)                    # sum messages of neighbours
        z1 = side @ Wgc + bgc
        bi = ego * side
        z2 = bi @ Wbi + bbi
        h = _leaky(z1) + _leaky(z2)
        m = None if masks is None or masks[k] is None else masks[k].astype(e0.dtype)
        hd = h if m is None else h * m / dt(keep)
        sq = (hd * hd).sum(1, keepdims=True)
        inv = 1.0 / np.sqrt(np.maximum(sq, dt(L2NORM_EPS)))
        outs.append((hd * inv).astype(e0.dtype))
        cache.append((ego, side, z1, z2, m, hd, sq, inv))
        ego = hd
    return np.concatenate(outs, axis=1), cache


def ngcf_loss_and_grad(A, AT, e0, weights, num_users, users, pos, neg, reg, masks=None, keep=1.0):
    """NGCF.py:94-110: sum softplus(-(pos - neg)) + reg * l2_loss(u, i, j) on the CONCATENATED
    embeddings, back-propagated by hand through the normalise / dropout / leaky-relu / two GEMMs /
    SpMM of every layer.  -> (mf_loss, emb_loss, dE0, [(dWgc, dbgc, dWbi, dbbi)], all_embeddings)."""
    dt = e0.dtype.type
    allE, cache = ngcf_forward(A, e0, weights, masks, keep)
    ru, ri, rj = np.asarray(users), num_users + np.asarray(pos), num_users + np.asarray(neg)
    eu, ei, ej = allE[ru], allE[ri], allE[rj]
    x = (eu * ei).sum(1) - (eu * ej).sum(1)
    l = np.where(x >= 0, np.log1p(np.exp(-x)), -x + np.log1p(np.exp(x)))      # softplus(-x)
    g = (-1.0 / (1.0 + np.exp(x))).astype(e0.dtype)[:, None]
    reg = dt(reg)
    emb = reg * dt(0.5) * ((eu * eu).sum() + (ei * ei).sum() + (ej * ej).sum())
    G = np.zeros_like(allE)
    np.add.at(G, ru, g * (ei - ej) + reg * eu)
    np.add.at(G, ri, g * eu + reg * ei)
    np.add.at(G, rj, -g * eu + reg * ej)
    # split the gradient of the concatenation back onto the layers
    dims = [e0.shape[1]] + [w[0].shape[1] for w in weights]
    offs = np.cumsum([0] + dims)
    d_ego_next = np.zeros((e0.shape[0], dims[-1]), e0.dtype)      # gradient flowing into layer K's raw output
    grads = [None] * len(weights)
    for k in range(len(weights) - 1, -1, -1):
        Wgc, bgc, Wbi, bbi = weights[k]
        ego, side, z1, z2, m, hd, sq, inv = cache[k]
        gn = G[:, offs[k + 1]:offs[k + 2]]                          # d loss / d normalised output of layer k
        # y = hd * inv,  inv = max(sq, eps)^-1/2  (l2_normalize)
        dot = (gn * hd).sum(1, keepdims=True)
        live = (sq > dt(L2NORM_EPS)).astype(e0.dtype)
        dhd = gn * inv - live * hd * dot * inv ** 3
        dhd = dhd + d_ego_next                                       # the un-normalised hd also feeds layer k+1
        dh = dhd if m is None else dhd * m / dt(keep)
        dz1 = dh * np.where(z1 > 0, dt(1), dt(LEAKY_ALPHA))
        dz2 = dh * np.where(z2 > 0, dt(1), dt(LEAKY_ALPHA))
        dWgc, dbgc = side.T @ dz1, dz1.sum(0, keepdims=True)
        bi = ego * side
        dWbi, dbbi = bi.T @ dz2, dz2.sum(0, keepdims=True)
        dside = dz1 @ Wgc.T
        dbi = dz2 @ Wbi.T
        dside = dside + dbi * ego
        dego = dbi * side + np.asarray(AT @ dside, dtype=e0.dtype)
        grads[k] = (dWgc.astype(e0.dtype), dbgc.astype(e0.dtype), dWbi.astype(e0.dtype), dbbi.astype(e0.dtype))
        d_ego_next = dego
    dE0 = (G[:, :dims[0]] + d_ego_next).astype(e0.dtype)
    return l.sum(), emb, dE0, grads, allE


class NGCFTrainer:
    """NGCF.train_model (NGCF.py:125-141) on the restatement above: BPR-softplus step with TF-1.12
    Adam; every gradient is dense here (E_0 through concat + SpMM, the layer weights by
    construction), so all variables take the dense ApplyAdam formulas, like LightGCN's table.
    Message dropout is always on in the reference (NGCF.py:193); masks come from `rs` so that a
    GPU run can be fed the same ones."""

    def __init__(self, A, e0, weights, num_users, lr=1e-3, reg=0.0, keep=0.9, rs=None):
        self.A, self.AT = A, A.T.tocsr()
        self.AT.sort_indices()
        self.e0 = np.array(e0, f32)
        self.W = [tuple(np.array(w, f32) for w in ws) for ws in weights]
        self.num_users, self.lr, self.reg, self.keep, self.t = num_users, lr, reg, keep, 0
        self.rs = rs or np.random.RandomState(0)
        z = np.zeros_like
        self.m_e, self.v_e = z(self.e0), z(self.e0)
        self.m_w = [tuple(z(w) for w in ws) for ws in self.W]
        self.v_w = [tuple(z(w) for w in ws) for ws in self.W]

    def draw_masks(self):
        if self.keep >= 1.0:
            return None
        n = self.e0.shape[0]
        return [(self.rs.rand(n, ws[0].shape[1]) < self.keep).astype(f32) for ws in self.W]

    def step(self, users, pos, neg, masks="draw"):
        masks = self.draw_masks() if isinstance(masks, str) else masks
        mf, emb, dE0, grads, _ = ngcf_loss_and_grad(self.A, self.AT, self.e0, self.W, self.num_users, users, pos, neg,
                                                    self.reg, masks, self.keep)
        hyper = [adam_lr_t(self.lr, 1, start_step=self.t)[0], 0.9, 0.999, 1e-8]
        opt_apply("adam", self.e0, dE0.astype(f32), self.m_e, self.v_e, None, hyper, dense_var=True)
        for k in range(len(self.W)):
            for j in range(4):
                opt_apply("adam", self.W[k][j], grads[k][j].astype(f32), self.m_w[k][j], self.v_w[k][j], None, hyper,
                          dense_var=True)
        self.t += 1
        return f32(mf), f32(emb)

    def embeddings(self, masks="draw"):
        """What `evaluate` reads (NGCF.py:144-146): the concatenated embeddings of one more forward
        pass -- with dropout still applied, as in the reference."""
        masks = self.draw_masks() if isinstance(masks, str) else masks
        allE, _ = ngcf_forward(self.A, self.e0, self.W, masks, self.keep)
        return allE[:self.num_users], allE[self.num_users:]
