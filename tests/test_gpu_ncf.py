"""GPU parity tests of the NCF family kernels (MLP.py / NeuMF.py) against oracle/tf_math.py.
fp32 with different summation orders (warp FMA chains vs numpy matmul) and atomics =>
tolerances stated per assertion."""
import numpy as np
import pytest
import torch

from oracle import tf_math

pytestmark = pytest.mark.gpu
KEYS = ("mf_user", "mf_item", "mlp_user", "mlp_item", "dense")


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_params(nu, ni, mf_dim, layers, n_towers, seed, scale=0.1):
    rs = np.random.RandomState(seed)
    mlp_dim = int(layers[0] / 2) if layers else 0
    P = {"mf_user": (rs.randn(nu, mf_dim) * scale).astype(np.float32) if mf_dim else None,
         "mf_item": (rs.randn(ni, mf_dim) * scale).astype(np.float32) if mf_dim else None,
         "mlp_user": (rs.randn(nu, mlp_dim) * scale).astype(np.float32) if layers else None,
         "mlp_item": (rs.randn(ni, mlp_dim) * scale).astype(np.float32) if layers else None,
         "dense": tf_math.ncf_init_dense(mlp_dim, layers, n_towers, rs) if layers else None}
    if layers:  # non-zero biases so that the bias path is exercised
        P["dense"] += (rs.randn(P["dense"].size) * 0.01).astype(np.float32)
    return P, mlp_dim


CASES = [
    # (mf_dim, layers, pairwise, loss, n_towers)   -- NeuMF pointwise is conf/NeuMF.properties
    (32, [64, 32, 16], False, "cross_entropy", 1),
    (16, [64, 32, 16], False, "square", 1),
    (32, [64, 32, 16], True, "bpr", 2),        # NeuMF pairwise: two towers
    (0, [64, 32, 16], True, "bpr", 1),         # MLP.py pairwise: shared tower
    (0, [32, 16], False, "cross_entropy", 1),  # MLP pointwise, 2 layers
    (8, [20, 10, 6, 4], True, "hinge", 2),     # odd widths, 4 layers
]


@pytest.mark.parametrize("mf_dim,layers,pairwise,loss,n_towers", CASES)
def test_ncf_grad_vs_oracle(mf_dim, layers, pairwise, loss, n_towers):
    from neurec_b200 import ops
    nu, ni, bs = 60, 90, 203            # ragged batch: last group of 8 is partial
    P, mlp_dim = make_params(nu, ni, mf_dim, layers, n_towers, 1)
    rs = np.random.RandomState(2)
    users = rs.randint(0, nu, bs).astype(np.int32)
    items = rs.randint(0, ni, bs).astype(np.int32)
    third = rs.randint(0, ni, bs).astype(np.int32) if pairwise else (rs.rand(bs) < 0.3).astype(np.float32)
    l, G, tU, tI = tf_math.ncf_grad(P, users, items, third, pairwise, loss, 0.01, 0.02, mlp_dim, layers, n_towers)
    shape = ops.NcfShape.make(nu, ni, mf_dim, layers, n_towers)
    assert shape.dense_size() == (P["dense"].size if layers else 0)
    dP = {k: dev(v) for k, v in P.items()}
    dG = {k: (torch.zeros_like(v) if v is not None else None) for k, v in dP.items()}
    dtU = torch.zeros(nu, dtype=torch.int32, device="cuda"); dtI = torch.zeros(ni, dtype=torch.int32, device="cuda")
    dl = torch.zeros(1, device="cuda")
    ops.ncf_grad(shape, dP, dev(users), dev(items), dev(third), pairwise, loss, 0.01, 0.02, dG, dtU, dtI, 3, dl)
    assert np.isclose(dl.item(), l, rtol=2e-5)
    for k in KEYS:
        if P[k] is None:
            continue
        # sums of <= ~400 fp32 terms of mixed sign in a different order: 1e-4 relative plus an
        # absolute term scaled by the largest gradient entry (near-cancelling sums)
        atol = 2e-6 * max(1.0, float(np.abs(G[k]).max()))
        assert np.allclose(dG[k].cpu().numpy(), G[k], rtol=1e-4, atol=atol), k
    assert np.array_equal(dtU.cpu().numpy() == 3, tU) and np.array_equal(dtI.cpu().numpy() == 3, tI)


def test_ncf_scores_vs_oracle():
    from neurec_b200 import ops
    nu, ni = 50, 333
    P, mlp_dim = make_params(nu, ni, 32, [64, 32, 16], 1, 5, scale=0.3)
    users = np.array([3, 7, 7, 49, 0], dtype=np.int32)
    shape = ops.NcfShape.make(nu, ni, 32, [64, 32, 16], 1)
    got = ops.ncf_scores(shape, {k: dev(v) for k, v in P.items()}, dev(users)).cpu().numpy()
    want = np.stack([tf_math.ncf_predict(P, np.full(ni, u), np.arange(ni), mlp_dim, [64, 32, 16]) for u in users])
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mf_dim,nu,ni", [(8, 943, 1682), (32, 130, 257), (64, 9, 128), (0, 37, 500)])
def test_ncf_scores_tile_kernel_equals_the_generic_kernel(mf_dim, nu, ni, monkeypatch):
    """The register-blocked predict kernel of the default tower (units of 4 users x 128 items, ragged last
    group / tile, GMF widths 0..64) against the generic warp-per-pair kernel on the same parameters."""
    from neurec_b200 import ops
    P, mlp_dim = make_params(nu, ni, mf_dim, [64, 32, 16], 1, 11 + mf_dim, scale=0.3)
    shape = ops.NcfShape.make(nu, ni, mf_dim, [64, 32, 16], 1)
    dP = {k: (dev(v) if v is not None else None) for k, v in P.items()}
    users = dev(np.random.RandomState(3).permutation(nu).astype(np.int32))
    fast = ops.ncf_scores(shape, dP, users).cpu().numpy()
    monkeypatch.setenv("NRC_NCF_SCORES_GENERIC", "1")
    slow = ops.ncf_scores(shape, dP, users).cpu().numpy()
    assert np.allclose(fast, slow, rtol=1e-5, atol=1e-6)
    assert np.abs(fast).max() > 1e-3


@pytest.mark.parametrize("pairwise,loss,opt", [(False, "cross_entropy", "adam"), (True, "bpr", "adam"),
                                               (False, "square", "rmsprop")])
def test_ncf_train_epoch_vs_oracle(ml100k, pairwise, loss, opt):
    """conf/NeuMF.properties shapes (embedding 32 per BASELINE configs[1], layers [64,32,16],
    bs 256) on real ml-100k ids, 12 steps.  Tolerance: 3e-5 absolute on every parameter,
    1e-4 relative on the per-step loss."""
    from neurec_b200 import ops
    d = ml100k
    nu, ni, bs, steps = d["num_users"], d["num_items"], 256, 12
    layers, mf_dim, nt = [64, 32, 16], 32, (2 if pairwise else 1)
    P, mlp_dim = make_params(nu, ni, mf_dim, layers, nt, 7, scale=0.01)
    rs = np.random.RandomState(8)
    all_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    perm = rs.permutation(len(all_users))[:bs * steps - 57]
    users, items = all_users[perm], d["train_indices"][perm]
    third = rs.randint(0, ni, len(users)).astype(np.int32) if pairwise else (rs.rand(len(users)) < 0.2).astype(np.float32)
    lr = 1e-3
    tr = tf_math.NCFTrainer(P, mlp_dim, layers, nt, opt, lr, loss, 0.0, 0.0, pairwise)
    want_loss = tr.epoch(users, items, third, bs)

    shape = ops.NcfShape.make(nu, ni, mf_dim, layers, nt)
    dP = {k: dev(v) for k, v in P.items()}
    i0, i1 = tf_math.SLOT_INIT[opt]
    G = {k: torch.zeros_like(v) for k, v in dP.items()}
    S0 = {k: torch.full_like(v, i0) for k, v in dP.items()}
    S1 = {k: torch.full_like(v, i1) for k, v in dP.items()}
    tU = torch.zeros(nu, dtype=torch.int32, device="cuda"); tI = torch.zeros(ni, dtype=torch.int32, device="cuda")
    step_loss = torch.zeros(steps, device="cuda")
    hyper = tf_math.DEFAULT_HYPER[opt](lr)
    lr_t = tf_math.adam_lr_t(lr, steps) if opt == "adam" else np.full(steps, lr, np.float32)
    n = ops.ncf_train_epoch(shape, dP, dev(users), dev(items), dev(third), bs, pairwise, loss, 0.0, 0.0, opt,
                            lr_t, hyper, G, S0, S1, tU, tI, 1, step_loss)
    assert n == steps
    assert np.allclose(step_loss.cpu().numpy(), want_loss, rtol=1e-4)
    for k in KEYS:
        assert np.abs(dP[k].cpu().numpy() - tr.P[k]).max() < 3e-5, k
    assert np.abs(tr.P["dense"] - P["dense"]).max() > 1e-4
