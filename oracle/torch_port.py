"""TEST / BASELINE INFRASTRUCTURE ONLY -- multi-threaded CPU restatement of the reference's
TensorFlow-1.12 training steps for TIMING (bench.py's ``cpu_baseline`` leg and ``--impl reference``
arm; BASELINE.md section 3: "fp32 restatement of the TF graph in torch-CPU with all host threads,
fed python lists per step exactly like feed_dict").  Never imported by the product.

Same formulas as oracle/tf_math.py (which is the PARITY oracle: numpy, operation-by-operation
rounding); this file trades the bit-level bookkeeping for speed -- vectorised gathers,
``index_add_`` for the IndexedSlices de-duplication, fused dense Adam -- so that the CPU arm is the
reference's algorithm at the speed a multi-threaded CPU framework reaches, not a strawman.  Each
function cites the reference lines it follows; tests/test_oracle.py checks it against tf_math.
"""
from __future__ import annotations

import numpy as np
import torch


def set_threads(n):
    torch.set_num_threads(max(1, int(n)))


class MFStep:
    """MF.py:54-76 + learner.py:2-41 + TF-1.12 Adam on IndexedSlices (dense over the whole table,
    adam.py::_apply_sparse_shared) or plain gradient descent (scatter_sub on the touched rows)."""

    def __init__(self, U, V, learner="adam", lr=1e-3, loss="bpr", reg=0.0, pairwise=True):
        self.U, self.V = torch.from_numpy(np.array(U, np.float32)), torch.from_numpy(np.array(V, np.float32))
        self.learner, self.lr, self.loss, self.reg, self.pairwise = learner, float(lr), loss, float(reg), pairwise
        if learner == "adam":
            z = torch.zeros_like
            self.mU, self.vU, self.mV, self.vV = z(self.U), z(self.U), z(self.V), z(self.V)
            self.p1, self.p2 = 0.9, 0.999
        elif learner != "gd":
            raise ValueError("torch_port times adam and gd only")

    def step(self, users, items, third):
        """One sess.run((loss, optimizer), feed_dict) of MF.py:97-101; inputs are python lists."""
        u = torch.as_tensor(users, dtype=torch.int64)
        i = torch.as_tensor(items, dtype=torch.int64)
        pu, qi = self.U[u], self.V[i]                                   # embedding_lookup, MF.py:57-58
        if self.pairwise:
            j = torch.as_tensor(third, dtype=torch.int64)
            qj = self.V[j]
            x = (pu * qi).sum(1) - (pu * qj).sum(1)                     # MF.py:59,66
            if self.loss == "bpr":                                      # learner.py:21-22
                loss = torch.nn.functional.softplus(-x).sum()
                g = -torch.sigmoid(-x)
            elif self.loss == "hinge":                                  # learner.py:23-24 [sic]
                loss = torch.clamp(x + 1, min=0).sum(); g = (x + 1 > 0).float()
            else:                                                       # learner.py:25-26
                loss = ((1 - x) ** 2).sum(); g = -2 * (1 - x)
            g = g[:, None]
            gU = g * (qi - qj) + self.reg * pu
            gi, gj = g * pu + self.reg * qi, -g * pu + self.reg * qj
            rows_v, grads_v = torch.cat([i, j]), torch.cat([gi, gj])
            if self.reg:                                                # MF.py:67 reg * l2_loss(p1, q2, q1)
                loss = loss + self.reg * 0.5 * ((pu * pu).sum() + (qi * qi).sum() + (qj * qj).sum())
        else:
            z = torch.as_tensor(third, dtype=torch.float32)
            x = (pu * qi).sum(1)
            if self.loss == "cross_entropy":                            # learner.py:33-34 (mean over the batch)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(x, z)
                g = (torch.sigmoid(x) - z) / len(x)
            else:                                                       # learner.py:37-38
                loss = ((z - x) ** 2).sum(); g = -2 * (z - x)
            g = g[:, None]
            gU, rows_v, grads_v = g * qi + self.reg * pu, i, g * pu + self.reg * qi
            if self.reg:                                                # MF.py:72 reg * l2_loss(p1, q1)
                loss = loss + self.reg * 0.5 * ((pu * pu).sum() + (qi * qi).sum())
        if self.learner == "gd":
            self.U.index_add_(0, u, gU, alpha=-self.lr)                 # scatter_sub of the IndexedSlices
            self.V.index_add_(0, rows_v, grads_v, alpha=-self.lr)
            return float(loss)
        dU = torch.zeros_like(self.U).index_add_(0, u, gU)              # _deduplicate_indexed_slices
        dV = torch.zeros_like(self.V).index_add_(0, rows_v, grads_v)
        lr_t = self.lr * np.sqrt(1 - self.p2) / (1 - self.p1)
        self.p1 *= 0.9; self.p2 *= 0.999
        for var, grad, m, v in ((self.U, dU, self.mU, self.vU), (self.V, dV, self.mV, self.vV)):
            m.mul_(0.9).add_(grad, alpha=0.1)                           # the WHOLE table, every step
            v.mul_(0.999).addcmul_(grad, grad, value=0.001)
            var.addcdiv_(m, v.sqrt().add_(1e-8), value=-lr_t)
        return float(loss)


class LightGCNStep:
    """LightGCN.py:132-166 + :130: forward propagation, BPR + reg on the layer-0 rows, backward
    through the same (symmetric) adjacency, dense Adam over E_0."""

    def __init__(self, A_csr, e0, num_users, n_layers, lr=0.01, reg=1e-3):
        A = A_csr.tocsr().astype(np.float32)
        self.A = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)),
                                         torch.from_numpy(A.indices.astype(np.int64)),
                                         torch.from_numpy(A.data), size=A.shape)
        self.e0 = torch.from_numpy(np.array(e0, np.float32))
        self.m, self.v = torch.zeros_like(self.e0), torch.zeros_like(self.e0)
        self.nu, self.L, self.lr, self.reg = num_users, n_layers, lr, reg
        self.p1, self.p2 = 0.9, 0.999

    def step(self, users, pos, neg):
        u = torch.as_tensor(users, dtype=torch.int64)
        i = torch.as_tensor(pos, dtype=torch.int64) + self.nu
        j = torch.as_tensor(neg, dtype=torch.int64) + self.nu
        e, acc = self.e0, self.e0.clone()
        for _ in range(self.L):                                         # LightGCN.py:139-147
            e = torch.sparse.mm(self.A, e)
            acc += e
        E = acc / (self.L + 1)
        pu, qi, qj = E[u], E[i], E[j]
        x = (pu * qi).sum(1) - (pu * qj).sum(1)
        loss = torch.nn.functional.softplus(-x).sum()
        g = (-torch.sigmoid(-x))[:, None] / (self.L + 1)
        G = torch.zeros_like(E)
        G.index_add_(0, u, g * (qi - qj)); G.index_add_(0, i, g * pu); G.index_add_(0, j, -g * pu)
        t = G
        for _ in range(self.L):                                         # backward: t_k = g + A^T t_{k+1}
            t = G + torch.sparse.mm(self.A, t)
        grad = t
        for rows in (u, i, j):                                          # LightGCN.py:161-164 regulariser
            grad.index_add_(0, rows, self.e0[rows], alpha=self.reg)
        lr_t = self.lr * np.sqrt(1 - self.p2) / (1 - self.p1)
        self.p1 *= 0.9; self.p2 *= 0.999
        self.m.mul_(0.9).add_(grad, alpha=0.1)
        self.v.mul_(0.999).addcmul_(grad, grad, value=0.001)
        self.e0.addcdiv_(self.m, self.v.sqrt().add_(1e-8), value=-lr_t)
        return float(loss)
