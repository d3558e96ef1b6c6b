"""GPU parity tests: CSR SpMM (bit-exact vs scipy's sequential CSR product = the TF CPU kernel's
order) and the LightGCN step (tolerance: atomics in the batch gradient) against oracle/tf_math."""
import numpy as np
import pytest
import torch

from oracle import tf_math

pytestmark = pytest.mark.gpu


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def csr_dev(A):
    return dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data.astype(np.float32))


def degree_order(A):
    return dev(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))


@pytest.fixture
def exact_spmm():
    """Sequential accumulation order (bit-identical to scipy / TF's CPU kernel) for the test's duration."""
    from neurec_b200 import ops
    ops.spmm_set_exact(True)
    yield
    ops.spmm_set_exact(False)


@pytest.mark.parametrize("dim", [64, 32, 128, 16, 50])
@pytest.mark.parametrize("adj_type", ["pre", "norm"])
def test_spmm_bit_exact_vs_scipy(ml100k, dim, adj_type, exact_spmm):
    from neurec_b200 import ops
    d = ml100k
    A = tf_math.lightgcn_adj(d["train_indptr"], d["train_indices"], d["num_users"], d["num_items"], adj_type)
    n = A.shape[0]
    X = (np.random.RandomState(dim).randn(n, dim)).astype(np.float32)
    want = (A @ X).astype(np.float32)
    ip, ix, va = csr_dev(A)
    got = ops.spmm_csr(ip, ix, va, dev(X))
    assert np.array_equal(got.cpu().numpy(), want)
    got2 = ops.spmm_csr(ip, ix, va, dev(X), row_order=degree_order(A))
    assert np.array_equal(got2.cpu().numpy(), want)
    # fused epilogue: y = bias + A.x ; sum = (sum + y) / div
    B = np.random.RandomState(1).randn(n, dim).astype(np.float32)
    S = np.random.RandomState(2).randn(n, dim).astype(np.float32)
    dS = dev(S)
    y = ops.spmm_csr(ip, ix, va, dev(X), bias=dev(B), sum_=dS, div=4.0)
    wy = (B + want).astype(np.float32)
    assert np.array_equal(y.cpu().numpy(), wy)
    assert np.array_equal(dS.cpu().numpy(), ((S + wy).astype(np.float32) / np.float32(4)).astype(np.float32))


def test_spmm_empty_rows_and_long_rows(exact_spmm):
    import scipy.sparse as sp
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    n, dim = 300, 64
    dense = (rs.rand(n, n) < 0.02) * rs.randn(n, n)
    dense[5] = 0; dense[17] = rs.randn(n)          # an empty row and a full row (> 32 nnz)
    A = sp.csr_matrix(dense.astype(np.float32)); A.sort_indices()
    X = rs.randn(n, dim).astype(np.float32)
    got = ops.spmm_csr(*csr_dev(A), dev(X))
    assert np.array_equal(got.cpu().numpy(), (A @ X).astype(np.float32))


@pytest.mark.parametrize("n_layers", [1, 3])
def test_lightgcn_propagate_bit_exact(ml100k, n_layers, exact_spmm):
    from neurec_b200 import ops
    d = ml100k
    A = tf_math.lightgcn_adj(d["train_indptr"], d["train_indices"], d["num_users"], d["num_items"], "pre")
    e0 = (np.random.RandomState(3).randn(A.shape[0], 64) * 0.1).astype(np.float32)
    want, _ = tf_math.lightgcn_propagate(A, e0, n_layers)
    ip, ix, va = csr_dev(A)
    got = ops.lightgcn_propagate(ip, ix, va, degree_order(A), dev(e0), n_layers)
    assert np.array_equal(got.cpu().numpy(), want)


def test_lightgcn_train_epoch_vs_oracle(ml100k):
    """conf/LightGCN.properties (lr 0.01, reg 1e-3, d 64, bs 1024) with n_layers=3 on the
    ml-100k graph, 6 steps on the same triplets.  Tolerance: 5e-5 abs on E0 (Adam with lr 1e-2
    amplifies the atomics' re-association at the first steps), losses 1e-4 rel."""
    from neurec_b200 import ops
    d = ml100k
    nu, ni, dim, L, bs, steps = d["num_users"], d["num_items"], 64, 3, 1024, 6
    A = tf_math.lightgcn_adj(d["train_indptr"], d["train_indices"], nu, ni, "pre")
    assert abs(A - A.T).max() < 1e-7          # 'pre' is symmetric: backward reuses the CSR
    rs = np.random.RandomState(4)
    lim = np.sqrt(6.0 / (nu + dim))
    e0 = rs.uniform(-lim, lim, (nu + ni, dim)).astype(np.float32)   # xavier-uniform-like
    all_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    perm = rs.permutation(len(all_users))[:bs * steps - 300]
    users, pos = all_users[perm], d["train_indices"][perm]
    neg = rs.randint(0, ni, len(users)).astype(np.int32)
    tr = tf_math.LightGCNTrainer(A, e0, nu, L, 0.01, 1e-3)
    want = tr.epoch(users, pos, neg, bs)

    de0 = dev(e0)
    z = lambda: torch.zeros_like(de0)
    m, v, ef, gf, ge, wa, wb = z(), z(), z(), z(), z(), z(), z()
    sl = torch.zeros(steps, 2, device="cuda")
    n = ops.lightgcn_train_epoch(csr_dev(A), None, degree_order(A), nu, ni, L, de0, m, v, dev(users),
                                 dev(pos), dev(neg), bs, 1e-3, tf_math.adam_lr_t(0.01, steps),
                                 [0.01, 0.9, 0.999, 1e-8], ef, gf, ge, (wa, wb), sl)
    assert n == steps
    assert np.allclose(sl.cpu().numpy(), want, rtol=1e-4)
    assert np.abs(de0.cpu().numpy() - tr.e0).max() < 5e-5
    assert float(gf.abs().max()) == 0.0 and float(ge.abs().max()) == 0.0
    # explicit transposed CSR gives the same result as the symmetric shortcut
    de0b = dev(e0)
    m, v, ef, gf, ge = z(), z(), z(), z(), z()
    AT = A.T.tocsr(); AT.sort_indices()
    ops.lightgcn_train_epoch(csr_dev(A), csr_dev(AT), None, nu, ni, L, de0b, m, v, dev(users), dev(pos),
                             dev(neg), bs, 1e-3, tf_math.adam_lr_t(0.01, steps), [0.01, 0.9, 0.999, 1e-8],
                             ef, gf, ge, (wa, wb), sl)
    assert np.abs(de0b.cpu().numpy() - tr.e0).max() < 5e-5


@pytest.mark.parametrize("dim", [32, 64, 128])
@pytest.mark.parametrize("adj_type", ["pre", "norm"])
def test_fast_spmm_matches_scipy_within_reassociation(ml100k, dim, adj_type):
    """The default accumulation order (several non-zeros per load, FFMA partial sums, long rows split
    over the CTA): deterministic and within fp32 re-association of the exact product -- also with the
    fused epilogue, with and without the degree order, and on rows far beyond the long-row threshold."""
    from neurec_b200 import ops
    d = ml100k
    A = tf_math.lightgcn_adj(d["train_indptr"], d["train_indices"], d["num_users"], d["num_items"], adj_type)
    assert np.diff(A.indptr).max() > 500                 # ml-100k has rows of 580+ non-zeros: the CTA path runs
    n = A.shape[0]
    X = (np.random.RandomState(dim).randn(n, dim)).astype(np.float32)
    want = (A.astype(np.float64) @ X.astype(np.float64))
    scale = np.abs(A).astype(np.float64) @ np.abs(X).astype(np.float64) + 1e-30
    ip, ix, va = csr_dev(A)
    got = ops.spmm_csr(ip, ix, va, dev(X), row_order=degree_order(A)).cpu().numpy()
    assert (np.abs(got - want) / scale).max() < 4e-7
    got2 = ops.spmm_csr(ip, ix, va, dev(X), row_order=degree_order(A)).cpu().numpy()
    assert np.array_equal(got, got2)                     # deterministic
    got3 = ops.spmm_csr(ip, ix, va, dev(X)).cpu().numpy()    # natural row order: a long row may share its unit with other rows -> other split
    assert (np.abs(got3 - want) / scale).max() < 4e-7
    B = np.random.RandomState(1).randn(n, dim).astype(np.float32)
    S = np.random.RandomState(2).randn(n, dim).astype(np.float32)
    dS = dev(S)
    y = ops.spmm_csr(ip, ix, va, dev(X), bias=dev(B), sum_=dS, div=4.0).cpu().numpy()
    assert np.abs(y - (B + want)).max() < 1e-5
    assert np.abs(dS.cpu().numpy() - (S + B + want) / 4.0).max() < 1e-5


def test_fast_spmm_empty_rows_tails_and_one_huge_row():
    import scipy.sparse as sp
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    n, dim = 700, 64
    dense = (rs.rand(n, n) < 0.03) * rs.randn(n, n)
    dense[5] = 0; dense[17] = rs.randn(n); dense[n - 1] = 0; dense[n - 1, 3] = 2.0
    A = sp.csr_matrix(dense.astype(np.float32)); A.sort_indices()
    X = rs.randn(n, dim).astype(np.float32)
    want = A.astype(np.float64) @ X.astype(np.float64)
    for order in (None, degree_order(A)):
        got = ops.spmm_csr(*csr_dev(A), dev(X), row_order=order).cpu().numpy()
        assert np.abs(got - want).max() < 2e-5
        assert (got[5] == 0).all() and np.allclose(got[n - 1], 2.0 * X[3])
