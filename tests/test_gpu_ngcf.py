"""GPU parity of the NGCF dense layers (csrc/ngcf.cu) around the CSR SpMM against
oracle/tf_math.py::ngcf_forward / ngcf_loss_and_grad (finite-difference pinned restatement of
NGCF.py:160-202 + 94-110; parity unpinned at the TensorFlow boundary), with explicit dropout masks
shared between the two sides, on the reference's own 'norm' adjacency D^-1 (A + I)."""
import numpy as np
import pytest
import torch

from oracle import tf_math

pytestmark = pytest.mark.gpu


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def csr_dev(A):
    A = A.tocsr().astype(np.float32)
    A.sort_indices()
    order = dev(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
    return (dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data)), order


def pack(weights):
    return np.concatenate([np.concatenate([w.reshape(-1) for w in ws]) for ws in weights]).astype(np.float32)


def _setup(d, emb, layers, keep, seed):
    nu, ni = d["num_users"], d["num_items"]
    A = tf_math.ngcf_adj(d["train_indptr"], d["train_indices"], nu, ni, "norm")
    rs = np.random.RandomState(seed)
    e0 = (rs.randn(nu + ni, emb) * 0.1).astype(np.float32)
    W = tf_math.ngcf_init_weights(rs, emb, layers)
    masks = None if keep >= 1.0 else [(rs.rand(nu + ni, w[0].shape[1]) < keep).astype(np.float32) for w in W]
    return A, e0, W, masks


@pytest.mark.parametrize("emb,layers,keep", [(16, [16, 16], 0.9), (64, [64, 32, 16], 0.9), (24, [40], 1.0)])
def test_ngcf_forward_vs_oracle(ml100k, emb, layers, keep):
    from neurec_b200 import ops
    d = ml100k
    A, e0, W, masks = _setup(d, emb, layers, keep, 1)
    want, _ = tf_math.ngcf_forward(A, e0, W, masks, keep)
    shape = ops.NgcfShape.make(d["num_users"], d["num_items"], emb, layers)
    assert shape.weights_size() == pack(W).size and shape.d_total == want.shape[1]
    csr, order = csr_dev(A)
    dm = None if masks is None else dev(np.concatenate([m.reshape(-1) for m in masks]))
    got = ops.ngcf_forward(shape, csr, order, dev(e0), dev(pack(W)), dm, keep).cpu().numpy()
    assert np.abs(got - want).max() < 2e-5


@pytest.mark.parametrize("emb,layers,keep,reg", [(16, [16, 16], 0.9, 1e-3), (64, [64, 32, 16], 0.9, 0.0), (24, [40], 1.0, 1e-2)])
def test_ngcf_gradients_vs_oracle(ml100k, emb, layers, keep, reg):
    from neurec_b200 import ops
    d = ml100k
    nu, ni = d["num_users"], d["num_items"]
    A, e0, W, masks = _setup(d, emb, layers, keep, 2)
    AT = A.T.tocsr(); AT.sort_indices()
    rs = np.random.RandomState(3)
    bs = 512
    users = rs.randint(0, nu, bs).astype(np.int32)
    pos = rs.randint(0, ni, bs).astype(np.int32); neg = rs.randint(0, ni, bs).astype(np.int32)
    mf, emb_l, dE0, grads, _ = tf_math.ngcf_loss_and_grad(A, AT, e0, W, nu, users, pos, neg, reg, masks, keep)
    shape = ops.NgcfShape.make(nu, ni, emb, layers)
    csr, order = csr_dev(A)
    tcsr, torder = csr_dev(AT)
    dm = None if masks is None else dev(np.concatenate([m.reshape(-1) for m in masks]))
    N, dt = shape.n_nodes, shape.d_total
    all_emb = torch.empty((N, dt), device="cuda"); G = torch.zeros((N, dt), device="cuda")
    gE = torch.empty((N, emb), device="cuda"); gW = torch.empty(shape.weights_size(), device="cuda")
    work = torch.empty(shape.work_floats(), device="cuda")
    loss2 = torch.zeros(2, device="cuda")
    ops.ngcf_grad(shape, csr, order, tcsr, torder, dev(e0), dev(pack(W)), dm, keep, dev(users), dev(pos), dev(neg), reg,
                  all_emb, G, gE, gW, work, loss2)
    l = loss2.cpu().numpy()
    assert abs(l[0] - mf) < 1e-4 * abs(mf) and abs(l[1] - emb_l) < 1e-4 * abs(emb_l) + 1e-7
    assert float(G.abs().max()) == 0.0                                   # accumulator left clean
    scale = max(1e-6, float(np.abs(dE0).max()))
    assert np.abs(gE.cpu().numpy() - dE0).max() < 2e-4 * scale + 1e-7
    want_w = pack(grads)
    got_w = gW.cpu().numpy()
    assert np.abs(got_w - want_w).max() < 2e-4 * max(1e-6, float(np.abs(want_w).max())) + 1e-7


def test_ngcf_training_steps_vs_oracle(ml100k):
    """conf/NGCF.properties (d 16, layers [16,16], bs 512, adam 1e-3, mess_dropout 0.1): 5 steps with the
    SAME dropout masks on both sides (drawn by the device generator, copied to the oracle)."""
    from neurec_b200 import ops
    d = ml100k
    nu, ni, emb, layers, keep, bs = d["num_users"], d["num_items"], 16, [16, 16], 0.9, 512
    A, e0, W, _ = _setup(d, emb, layers, 1.0, 5)
    shape = ops.NgcfShape.make(nu, ni, emb, layers)
    csr, order = csr_dev(A)
    AT = A.T.tocsr(); AT.sort_indices()
    tcsr, torder = csr_dev(AT)
    tr = tf_math.NGCFTrainer(A, e0, W, nu, 1e-3, 1e-4, keep)
    dE, dW = dev(e0), dev(pack(W))
    N, dt = shape.n_nodes, shape.d_total
    z = torch.zeros
    all_emb, G = z((N, dt), device="cuda"), z((N, dt), device="cuda")
    gE, gW = torch.zeros_like(dE), torch.zeros_like(dW)
    mE, vE, mW, vW = torch.zeros_like(dE), torch.zeros_like(dE), torch.zeros_like(dW), torch.zeros_like(dW)
    work = torch.empty(shape.work_floats(), device="cuda")
    masks = torch.empty(shape.mask_floats(), device="cuda")
    rs = np.random.RandomState(9)
    lr_t = tf_math.adam_lr_t(1e-3, 5)
    for s in range(5):
        users = rs.randint(0, nu, bs).astype(np.int32); pos = rs.randint(0, ni, bs).astype(np.int32); neg = rs.randint(0, ni, bs).astype(np.int32)
        ops.dropout_mask(masks.numel(), keep, 2017, s, out=masks)
        mh = masks.cpu().numpy()
        assert 0.88 < mh.mean() < 0.92 and set(np.unique(mh).tolist()) <= {0.0, 1.0}
        offs = np.cumsum([0] + [N * w for w in layers])
        host_masks = [mh[offs[k]:offs[k + 1]].reshape(N, layers[k]) for k in range(len(layers))]
        want = tr.step(users, pos, neg, masks=host_masks)
        loss2 = torch.zeros(2, device="cuda")
        ops.ngcf_grad(shape, csr, order, tcsr, torder, dE, dW, masks, keep, dev(users), dev(pos), dev(neg), 1e-4,
                      all_emb, G, gE, gW, work, loss2)
        ops.opt_apply_multi("adam", [(dE, gE, mE, vE, None, True), (dW, gW, mW, vW, None, True)], 0,
                            [float(lr_t[s]), 0.9, 0.999, 1e-8])
        got = loss2.cpu().numpy()
        assert abs(got[0] - want[0]) < 2e-4 * abs(want[0]), (s, got, want)
    assert np.abs(dE.cpu().numpy() - tr.e0).max() < 5e-5
    assert np.abs(dW.cpu().numpy() - pack(tr.W)).max() < 5e-5
