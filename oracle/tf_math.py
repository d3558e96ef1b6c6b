"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of the reference's TensorFlow-1.12 graphs.

Parity status: UNPINNED at the TensorFlow boundary.  tensorflow==1.12.3 (requirements.txt:4 of
the reference) is a third-party dependency that is neither vendored in /root/reference nor
installable here (no network, Python 3.12), and the reference ships no tests or golden values
for it.  The arithmetic below restates TF's published algorithms; every function cites the
reference call site it stands in for.  What IS pinned: the formulas agree with closed-form
gradients checked by finite differences in tests/test_oracle.py.

Restated call sites (paths relative to /root/reference):
  model/general_recommender/MF.py:54-76        BPRMF / pointwise "GMF" graph
  util/learner.py:2-41                         optimizer / pairwise_loss / pointwise_loss
  util/tool.py:216-224                         l2_loss, log_loss
  model/general_recommender/NeuMF.py:69-104    NeuMF graph (see neumf_*)
  model/general_recommender/MLP.py:45-87       MLP graph
  model/general_recommender/LightGCN.py:35-78,132-166   adjacency + propagation + loss
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
            s0[rows] = s0[rows] + (gg * gg - s0[rows]) * (one - h[1])
            s1[rows] = s1[rows] * h[2] + (h[0] * gg) * (one / np.sqrt(s0[rows] + h[3]))
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
