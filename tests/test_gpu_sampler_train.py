"""GPU parity tests of the negative sampler and of the MF-family training step.

Sampler: bit-exact against the CPU restatement of the product's Philox stream + the
reference's contract (data/sampler.py:71-90): never a train item of the user, uniform over the
rest, aligned with the positives, fresh per epoch.
Training: fp32 with atomics => tolerance (stated per test) against oracle/tf_math.py fed the
SAME triplets; the optimizer arithmetic alone is bit-exact given an identical gradient."""
import numpy as np
import pytest
import torch

import oracle
from oracle import tf_math

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------ sampler
def test_sampler_bit_exact_vs_cpu_restatement(ml100k):
    from neurec_b200 import ops
    d = ml100k
    users = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
    for neg_num, seed, sid, first in [(1, 2018, 0, 0), (4, 7, 3, 0), (2, 2 ** 40 + 5, 2 ** 33 + 1, 1000)]:
        got = ops.sample_negatives(dev(d["train_indptr"]), dev(d["train_indices"]), dev(users), neg_num,
                                   d["num_items"], seed, sid, first).cpu().numpy()
        want = oracle.philox_sample_negatives(d["train_indptr"], d["train_indices"], users, neg_num,
                                              d["num_items"], seed, sid, first)
        assert np.array_equal(got, want)


def test_sampler_contract(ml100k):
    from neurec_b200 import ops
    d = ml100k
    ip, ix = d["train_indptr"], d["train_indices"]
    users = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(ip))
    a = ops.sample_negatives(dev(ip), dev(ix), dev(users), 1, d["num_items"], 1, 0).cpu().numpy()[:, 0]
    b = ops.sample_negatives(dev(ip), dev(ix), dev(users), 1, d["num_items"], 1, 1).cpu().numpy()[:, 0]
    assert a.min() >= 0 and a.max() < d["num_items"]
    assert (a != b).mean() > 0.99                      # new negatives every epoch (stream_id)
    train_sets = [set(ix[ip[u]:ip[u + 1]].tolist()) for u in range(d["num_users"])]
    assert not any(int(n) in train_sets[u] for u, n in zip(users, a))
    # shard invariance: sampling the second half alone gives the same numbers
    h = len(users) // 2
    c = ops.sample_negatives(dev(ip), dev(ix), dev(users[h:]), 1, d["num_items"], 1, 0, first_index=h)
    assert np.array_equal(c.cpu().numpy()[:, 0], a[h:])
    # uniformity over the allowed items of the heaviest user (chi-square, 5 sigma)
    u = int(np.argmax(np.diff(ip)))
    reps = 200000
    s = ops.sample_negatives(dev(ip), dev(ix), dev(np.full(reps, u, np.int32)), 1, d["num_items"], 3, 9)
    s = s.cpu().numpy()[:, 0]
    allowed = np.setdiff1d(np.arange(d["num_items"]), ix[ip[u]:ip[u + 1]])
    cnt = np.bincount(s, minlength=d["num_items"])[allowed]
    exp = reps / len(allowed)
    chi2 = ((cnt - exp) ** 2 / exp).sum()
    dof = len(allowed) - 1
    assert abs(chi2 - dof) < 5 * np.sqrt(2 * dof)


def test_batch_randint_choice_modes():
    from neurec_b200 import ops
    rs = np.random.RandomState(0)
    high = 300
    sizes = rs.randint(1, 40, 57)
    optr = np.zeros(58, np.int64); optr[1:] = np.cumsum(sizes)
    ep, ei = oracle.lists_to_csr([rs.choice(high, rs.randint(0, 200), replace=False) for _ in range(57)])
    for replace in (True, False):
        got = ops.batch_randint_choice(high, dev(optr), int(optr[-1]), replace, dev(ep), dev(ei), 11, 2)
        got = got.cpu().numpy()
        want = oracle.philox_batch_choice(high, optr, replace, (ep, ei), 11, 2)
        assert np.array_equal(got, want)
        for r in range(57):
            row = got[optr[r]:optr[r + 1]]
            assert not np.isin(row, ei[ep[r]:ep[r + 1]]).any()
            if not replace:
                assert len(set(row.tolist())) == len(row)
    got = ops.batch_randint_choice(high, dev(optr), int(optr[-1]), True, None, None, 1, 0).cpu().numpy()
    assert got.min() >= 0 and got.max() < high
    with pytest.raises(ValueError):
        ops.sample_negatives(dev(ep), dev(ei), dev(np.zeros(3, np.int32)), 0, high, 1, 0)


def test_captured_step_graph_matches_oracle(ml100k):
    """The reference-facing per-batch path (host id arrays -> pinned block -> ONE graph launch ->
    loss back) must train exactly like the eager path: 8 BPR/Adam steps vs the numpy oracle,
    lr_t read from device memory inside the captured graph."""
    import ctypes
    from neurec_b200 import _lib, ops
    d = ml100k
    lib = _lib.load()
    nu, ni, dim, bs, steps = d["num_users"], d["num_items"], 64, 512, 8
    rs = np.random.RandomState(3)
    U0 = (rs.randn(nu, dim) * 0.01).astype(np.float32); V0 = (rs.randn(ni, dim) * 0.01).astype(np.float32)
    all_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    perm = rs.permutation(len(all_users))[:bs * steps]
    users, pos = all_users[perm].copy(), d["train_indices"][perm].copy()
    neg = rs.randint(0, ni, len(users)).astype(np.int32)
    tr = tf_math.MFTrainer(U0, V0, "adam", 1e-3, "bpr", 0.0, True)
    want = tr.epoch(users, pos, neg, bs)

    dU, dV = dev(U0), dev(V0)
    z = torch.zeros_like
    gU, gV, mU, vU, mV, vV = z(dU), z(dV), z(dU), z(dU), z(dV), z(dV)
    tU = torch.zeros(nu, dtype=torch.int32, device="cuda"); tV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    staging = torch.zeros(3 * bs + 4, dtype=torch.int32, device="cuda")
    pin = torch.zeros(3 * bs + 4, dtype=torch.int32).pin_memory()
    loss_pin = torch.zeros(4).pin_memory()
    step_loss = torch.zeros(4, device="cuda")
    s = torch.cuda.Stream()
    graph = ctypes.c_void_p()
    vp = ctypes.c_void_p
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        st = vp(s.cuda_stream)
        _lib.check(lib.nrc_graph_capture_begin(st))
        _lib.check(lib.nrc_graph_stage_async(vp(pin.data_ptr()), vp(staging.data_ptr()), (3 * bs + 1) * 4, st))
        _lib.check(lib.nrc_opt_set_lr_source(vp(staging.data_ptr() + 12 * bs)))
        ops.mf_train_epoch(dU, dV, staging[:bs], staging[bs:2 * bs], staging[2 * bs:3 * bs], bs, True, "bpr", 0.0,
                           "adam", np.zeros(1, np.float32), [0.0, 0.9, 0.999, 1e-8], gU, gV, tU, tV, mU, vU, mV,
                           vV, 1, step_loss)
        _lib.check(lib.nrc_opt_set_lr_source(None))
        _lib.check(lib.nrc_graph_fetch_async(vp(step_loss.data_ptr()), vp(loss_pin.data_ptr()), 1, st))
        _lib.check(lib.nrc_graph_capture_end(st, ctypes.byref(graph)))
    torch.cuda.synchronize()
    assert float(dU.cpu().numpy().std()) > 0 and np.array_equal(dU.cpu().numpy(), U0)   # capture ran nothing
    lr_t = tf_math.adam_lr_t(1e-3, steps)
    got = []
    for k in range(steps):
        o = k * bs
        _lib.check(lib.nrc_graph_step(graph, vp(users[o:].ctypes.data), vp(pos[o:].ctypes.data),
                                      vp(neg[o:].ctypes.data), bs, float(lr_t[k]), vp(pin.data_ptr()),
                                      vp(s.cuda_stream)))
        got.append(float(loss_pin[0]))
    lib.nrc_graph_destroy(graph)
    assert np.allclose(got, want, rtol=1e-4)
    assert np.abs(dU.cpu().numpy() - tr.U).max() < 2e-5 and np.abs(dV.cpu().numpy() - tr.V).max() < 2e-5


def test_graph_ring_burst_matches_oracle(ml100k):
    """nrc_graph_run_steps: 11 BPR/Adam steps from host arrays through a ring of 4 pinned blocks --
    two bursts of the 4-step burst graph (H2D chain and loss-D2H chain captured on side streams with
    nrc_graph_depend, so copies overlap the previous step's kernels) and 3 leftover steps through the
    single-step graphs.  Every step stages its own batch and returns its own loss.  Same result as
    the numpy oracle stepping through the same batches."""
    import ctypes
    from neurec_b200 import _lib, ops
    d = ml100k
    lib = _lib.load()
    nu, ni, dim, bs, steps, ring = d["num_users"], d["num_items"], 64, 512, 11, 4
    rs = np.random.RandomState(5)
    U0 = (rs.randn(nu, dim) * 0.01).astype(np.float32); V0 = (rs.randn(ni, dim) * 0.01).astype(np.float32)
    all_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    perm = rs.permutation(len(all_users))[:bs * steps]
    users, pos = all_users[perm].copy(), d["train_indices"][perm].copy()
    neg = rs.randint(0, ni, len(users)).astype(np.int32)
    tr = tf_math.MFTrainer(U0, V0, "adam", 1e-3, "bpr", 0.0, True)
    want = tr.epoch(users, pos, neg, bs)

    dU, dV = dev(U0), dev(V0)
    z = torch.zeros_like
    gU, gV, mU, vU, mV, vV = z(dU), z(dV), z(dU), z(dU), z(dV), z(dV)
    tU = torch.zeros(nu, dtype=torch.int32, device="cuda"); tV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    stag = torch.zeros((ring, 3 * bs + 4), dtype=torch.int32, device="cuda")
    loss_dev = torch.zeros((ring, 4), device="cuda")
    pins = [torch.zeros(3 * bs + 4, dtype=torch.int32).pin_memory() for _ in range(ring)]
    loss_pins = [torch.zeros(4).pin_memory() for _ in range(ring)]
    main, s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    vp = ctypes.c_void_p
    sp = lambda x: vp(x.cuda_stream)

    def capture_step(r, cin, cout):
        _lib.check(lib.nrc_graph_stage_async(vp(pins[r].data_ptr()), vp(stag[r].data_ptr()), (3 * bs + 1) * 4, sp(cin)))
        if cin is not main:
            _lib.check(lib.nrc_graph_depend(sp(cin), sp(main)))
        _lib.check(lib.nrc_opt_set_lr_source(vp(stag[r].data_ptr() + 12 * bs)))
        with torch.cuda.stream(main):
            ops.mf_train_epoch(dU, dV, stag[r, :bs], stag[r, bs:2 * bs], stag[r, 2 * bs:3 * bs], bs, True, "bpr",
                               0.0, "adam", np.zeros(1, np.float32), [0.0, 0.9, 0.999, 1e-8], gU, gV, tU, tV, mU,
                               vU, mV, vV, 1, loss_dev[r])
        _lib.check(lib.nrc_opt_set_lr_source(None))
        if cout is not main:
            _lib.check(lib.nrc_graph_depend(sp(main), sp(cout)))
        _lib.check(lib.nrc_graph_fetch_async(vp(loss_dev[r].data_ptr()), vp(loss_pins[r].data_ptr()), 1, sp(cout)))

    graphs = []
    torch.cuda.synchronize()
    for r in range(ring):
        g = ctypes.c_void_p()
        _lib.check(lib.nrc_graph_capture_begin(sp(main)))
        capture_step(r, main, main)
        _lib.check(lib.nrc_graph_capture_end(sp(main), ctypes.byref(g)))
        graphs.append(g)
    burst = ctypes.c_void_p()
    _lib.check(lib.nrc_graph_capture_begin(sp(main)))
    _lib.check(lib.nrc_graph_depend(sp(main), sp(s_in)))
    _lib.check(lib.nrc_graph_depend(sp(main), sp(s_out)))
    for r in range(ring):
        capture_step(r, s_in, s_out)
    _lib.check(lib.nrc_graph_depend(sp(s_in), sp(main)))
    _lib.check(lib.nrc_graph_depend(sp(s_out), sp(main)))
    _lib.check(lib.nrc_graph_capture_end(sp(main), ctypes.byref(burst)))
    torch.cuda.synchronize()
    assert np.array_equal(dU.cpu().numpy(), U0)                      # capturing ran nothing

    lr_t = np.ascontiguousarray(tf_math.adam_lr_t(1e-3, steps), dtype=np.float32)
    total = ctypes.c_double(0.0)
    c_graphs = (ctypes.c_void_p * ring)(*[g.value for g in graphs])
    c_pins = (ctypes.c_void_p * ring)(*[p.data_ptr() for p in pins])
    c_loss = (ctypes.c_void_p * ring)(*[p.data_ptr() for p in loss_pins])
    _lib.check(lib.nrc_graph_run_steps(burst, c_graphs, ring, vp(users.ctypes.data), vp(pos.ctypes.data),
                                       vp(neg.ctypes.data), bs, vp(lr_t.ctypes.data), steps, c_pins, c_loss, 1,
                                       ctypes.byref(total), sp(main)))
    for g in graphs + [burst]:
        lib.nrc_graph_destroy(g)
    assert np.isclose(total.value, float(np.sum(want, dtype=np.float64)), rtol=1e-4)
    assert np.isclose(float(loss_pins[(steps - 1) % ring][0]), want[-1], rtol=1e-4)   # last step's own loss
    assert np.abs(dU.cpu().numpy() - tr.U).max() < 2e-5 and np.abs(dV.cpu().numpy() - tr.V).max() < 2e-5


# ----------------------------------------------------------------------------- training
def _tables(nu, ni, dim, seed=0, scale=0.1):
    rs = np.random.RandomState(seed)
    return (rs.randn(nu, dim) * scale).astype(np.float32), (rs.randn(ni, dim) * scale).astype(np.float32)


@pytest.mark.parametrize("loss", ["bpr", "hinge", "square"])
@pytest.mark.parametrize("dim", [64, 20])
def test_pairwise_grad_vs_oracle(loss, dim):
    from neurec_b200 import ops
    nu, ni, bs = 50, 80, 512            # heavy duplication of users and items inside the batch
    U, V = _tables(nu, ni, dim, 1)
    rs = np.random.RandomState(2)
    users = rs.randint(0, nu, bs).astype(np.int32)
    pos = rs.randint(0, ni, bs).astype(np.int32)
    neg = rs.randint(0, ni, bs).astype(np.int32)
    l, gU, gV, tU, tV = tf_math.mf_pairwise_grad(U, V, users, pos, neg, loss, reg=0.01)
    dgU = torch.zeros(nu, dim, device="cuda"); dgV = torch.zeros(ni, dim, device="cuda")
    dtU = torch.zeros(nu, dtype=torch.int32, device="cuda"); dtV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    dl = torch.zeros(1, device="cuda")
    ops.mf_pairwise_grad(dev(U), dev(V), dev(users), dev(pos), dev(neg), loss, 0.01, dgU, dgV, dtU, dtV, 5, dl)
    # tolerance: fp32 sums of <= ~30 duplicate contributions in arbitrary (atomic) order
    assert np.allclose(dgU.cpu().numpy(), gU, rtol=1e-4, atol=1e-6)
    assert np.allclose(dgV.cpu().numpy(), gV, rtol=1e-4, atol=1e-6)
    assert np.isclose(dl.item(), l, rtol=1e-5)
    assert np.array_equal(dtU.cpu().numpy() == 5, tU) and np.array_equal(dtV.cpu().numpy() == 5, tV)


@pytest.mark.parametrize("loss", ["cross_entropy", "square"])
def test_pointwise_grad_vs_oracle(loss):
    from neurec_b200 import ops
    nu, ni, bs, dim = 40, 60, 256, 32
    U, V = _tables(nu, ni, dim, 3, scale=0.5)
    rs = np.random.RandomState(4)
    users = rs.randint(0, nu, bs).astype(np.int32)
    items = rs.randint(0, ni, bs).astype(np.int32)
    labels = (rs.rand(bs) < 0.2).astype(np.float32)
    l, gU, gV, tU, tV = tf_math.mf_pointwise_grad(U, V, users, items, labels, loss, reg=0.001)
    dgU = torch.zeros(nu, dim, device="cuda"); dgV = torch.zeros(ni, dim, device="cuda")
    dtU = torch.zeros(nu, dtype=torch.int32, device="cuda"); dtV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    dl = torch.zeros(1, device="cuda")
    ops.mf_pointwise_grad(dev(U), dev(V), dev(users), dev(items), dev(labels), loss, 0.001, dgU, dgV, dtU, dtV, 1, dl)
    assert np.allclose(dgU.cpu().numpy(), gU, rtol=1e-4, atol=1e-6)
    assert np.allclose(dgV.cpu().numpy(), gV, rtol=1e-4, atol=1e-6)
    assert np.isclose(dl.item(), l, rtol=1e-5)


@pytest.mark.parametrize("dim", [128, 64, 32])
def test_fused_single_pass_sgd_vs_oracle(dim):
    """Large-table path: with no repeated row inside the batch the one-pass kernel equals the
    TF gd step (1e-6 abs); with heavy repeats every contribution still lands (atomics), only
    the read/update interleaving differs -> compare against sequential per-triplet SGD bounds."""
    from neurec_b200 import ops
    rs = np.random.RandomState(dim)
    nu, ni, bs = 3000, 5000, 1024
    U, V = _tables(nu, ni, dim, 9, scale=0.3)
    users = rs.permutation(nu)[:bs].astype(np.int32)
    items = rs.permutation(ni)[:2 * bs].astype(np.int32)        # all rows distinct
    pos, neg = items[:bs], items[bs:]
    tr = tf_math.MFTrainer(U, V, "gd", 0.05, "bpr", 0.01, True)
    want_loss = tr.step(users, pos, neg)
    dU, dV, dl = dev(U), dev(V), torch.zeros(1, device="cuda")
    ops.mf_bpr_sgd_fused(dU, dV, dev(users), dev(pos), dev(neg), 0.05, 0.01, dl)
    assert np.isclose(dl.item(), want_loss, rtol=1e-5)
    assert np.abs(dU.cpu().numpy() - tr.U).max() < 1e-6 and np.abs(dV.cpu().numpy() - tr.V).max() < 1e-6
    # repeats: one hot item in every triplet -> its row receives all 1024 contributions
    pos2 = np.full(bs, 7, np.int32)
    tr2 = tf_math.MFTrainer(U, V, "gd", 1e-4, "bpr", 0.0, True)
    tr2.step(users, pos2, neg)
    dU, dV = dev(U), dev(V)
    ops.mf_bpr_sgd_fused(dU, dV, dev(users), dev(pos2), dev(neg), 1e-4, 0.0, dl)
    # hogwild-inside-batch deviation is second order in lr: |delta| <= ~lr^2 * bs * |grad|^2
    assert np.abs(dV.cpu().numpy()[7] - tr2.V[7]).max() < 3e-2 * np.abs(tr2.V[7] - V[7]).max() + 1e-6
    assert np.abs(dU.cpu().numpy() - tr2.U).max() < 1e-5
    with pytest.raises(_NrcLimit):
        U48, V48 = _tables(nu, ni, 48, 1)
        ops.mf_bpr_sgd_fused(dev(U48), dev(V48), dev(users), dev(pos), dev(neg), 0.1, 0.0, dl)


from neurec_b200._lib import NrcError as _NrcLimit  # noqa: E402


@pytest.mark.parametrize("opt", ["gd", "adam", "adagrad", "rmsprop", "momentum"])
def test_optimizer_apply_bit_exact(opt):
    """Given the same gradient the TF-1.12 update rules are bit-identical to the numpy oracle."""
    from neurec_b200 import ops
    rs = np.random.RandomState(5)
    rows, dim = 37, 24
    var = rs.randn(rows, dim).astype(np.float32)
    i0, i1 = tf_math.SLOT_INIT[opt]
    s0 = None if i0 is None else (np.abs(rs.randn(rows, dim)) * 0.1 + i0).astype(np.float32)
    s1 = None if i1 is None else (np.abs(rs.randn(rows, dim)) * 0.1 + i1).astype(np.float32)
    touched = rs.rand(rows) < 0.4
    g = np.zeros((rows, dim), np.float32)
    g[touched] = rs.randn(int(touched.sum()), dim).astype(np.float32)
    hyper = tf_math.DEFAULT_HYPER[opt](0.05)
    if opt == "adam":
        hyper[0] = float(tf_math.adam_lr_t(0.05, 3)[2])
    dvar, dg = dev(var), dev(g)
    ds0 = dev(s0) if s0 is not None else None
    ds1 = dev(s1) if s1 is not None else None
    stamps = dev(np.where(touched, 9, 3).astype(np.int32))
    ops.opt_apply_rows(opt, dvar, dg, ds0, ds1, stamps, 9, hyper)
    tf_math.opt_apply(opt, var, g, s0, s1, touched, hyper)
    assert np.array_equal(dvar.cpu().numpy(), var)
    if s0 is not None:
        assert np.array_equal(ds0.cpu().numpy(), s0)
    if s1 is not None:
        assert np.array_equal(ds1.cpu().numpy(), s1)
    assert float(dg.abs().max()) == 0.0  # gradient buffer zeroed for the next step


@pytest.mark.parametrize("opt,pairwise,loss", [("adam", True, "bpr"), ("gd", True, "bpr"),
                                               ("adagrad", False, "cross_entropy"),
                                               ("rmsprop", True, "hinge"), ("momentum", False, "square")])
def test_train_epoch_vs_oracle(ml100k, opt, pairwise, loss):
    """A real ml-100k epoch prefix (conf/MF.properties: bs 512, d 64): same triplet stream into
    the kernel and the oracle.  Tolerance: tables within 2e-5 abs after 20 steps (fp32
    re-association through atomics; Adam's division by sqrt(v)+1e-8 amplifies gradient ulps
    at the first steps), per-step loss within 1e-4 relative."""
    from neurec_b200 import ops
    d = ml100k
    nu, ni, dim, bs, steps = d["num_users"], d["num_items"], 64, 512, 20
    rs = np.random.RandomState(11)
    U0 = (rs.randn(nu, dim) * 0.01).astype(np.float32)
    V0 = (rs.randn(ni, dim) * 0.01).astype(np.float32)
    all_users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    perm = rs.permutation(len(all_users))[:bs * steps - 100]     # last batch is short
    users = all_users[perm]; items = d["train_indices"][perm]
    lr = 0.05 if opt in ("gd", "momentum") else 1e-3
    if opt == "adagrad":
        lr = 0.01
    if pairwise:
        neg = oracle.philox_sample_negatives(d["train_indptr"], d["train_indices"], users, 1, ni, 5, 0)[:, 0]
        third = neg
    else:
        third = (rs.rand(len(users)) < 0.5).astype(np.float32)
    tr = tf_math.MFTrainer(U0, V0, opt, lr, loss, reg=0.001, pairwise=pairwise)
    want_loss = tr.epoch(users, items, third, bs)

    dU, dV = dev(U0), dev(V0)
    z = lambda a: torch.zeros_like(a)
    i0, i1 = tf_math.SLOT_INIT[opt]
    mk = lambda a, v: None if v is None else torch.full_like(a, v)
    s0U, s1U, s0V, s1V = mk(dU, i0), mk(dU, i1), mk(dV, i0), mk(dV, i1)
    gU, gV = z(dU), z(dV)
    tU = torch.zeros(nu, dtype=torch.int32, device="cuda"); tV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    step_loss = torch.zeros(steps, device="cuda")
    hyper = tf_math.DEFAULT_HYPER[opt](lr)
    lr_t = tf_math.adam_lr_t(lr, steps) if opt == "adam" else np.full(steps, lr, np.float32)
    n = ops.mf_train_epoch(dU, dV, dev(users), dev(items), dev(third), bs, pairwise, loss, 0.001, opt,
                           lr_t, hyper, gU, gV, tU, tV, s0U, s1U, s0V, s1V, 1, step_loss)
    assert n == steps
    assert np.allclose(step_loss.cpu().numpy(), want_loss, rtol=1e-4)
    assert np.abs(dU.cpu().numpy() - tr.U).max() < 2e-5
    assert np.abs(dV.cpu().numpy() - tr.V).max() < 2e-5
    # the tables really moved
    assert np.abs(tr.U - U0).max() > 1e-4


def test_bprmf_epoch_ndcg_at_10_matches_the_cpu_path(ml100k):
    """BASELINE config 1 end to end: one full BPRMF epoch on ml-100k (157 steps of 512, Adam 1e-3,
    d=64) trained by the CUDA path and by the CPU oracle on the same triplet stream, then the
    full-catalogue evaluation of each model by its own evaluator (GPU kernels / C oracle).
    north_star's bar: NDCG@10 within 1e-5 of the CPU path.  (The tables differ by ~1e-7 through
    fp32 re-association in the atomics; a flipped near-tie at a hit position would move the mean
    by 1/943 x ~0.05, so this bound is met only while no such flip happens -- it is asserted.)"""
    from neurec_b200 import ops
    d = ml100k
    nu, ni, dim, bs = d["num_users"], d["num_items"], 64, 512
    rs = np.random.RandomState(2017)
    U0 = (rs.randn(nu, dim) * 0.01).astype(np.float32)
    V0 = (rs.randn(ni, dim) * 0.01).astype(np.float32)
    users = np.repeat(np.arange(nu, dtype=np.int32), np.diff(d["train_indptr"]))
    items = d["train_indices"]
    neg = oracle.philox_sample_negatives(d["train_indptr"], d["train_indices"], users, 1, ni, 2018, 0)[:, 0]
    perm = rs.permutation(len(users))
    users, items, neg = users[perm], items[perm], neg[perm]
    steps = (len(users) + bs - 1) // bs
    tr = tf_math.MFTrainer(U0, V0, "adam", 1e-3, "bpr", reg=0.0, pairwise=True)
    tr.epoch(users, items, neg, bs)

    dU, dV = dev(U0), dev(V0)
    z = lambda a: torch.zeros_like(a)
    step_loss = torch.zeros(steps, device="cuda")
    n = ops.mf_train_epoch(dU, dV, dev(users), dev(items), dev(neg), bs, True, "bpr", 0.0, "adam",
                           tf_math.adam_lr_t(1e-3, steps), [1e-3, 0.9, 0.999, 1e-8], z(dU), z(dV),
                           torch.zeros(nu, dtype=torch.int32, device="cuda"),
                           torch.zeros(ni, dtype=torch.int32, device="cuda"), z(dU), z(dU), z(dV), z(dV), 1,
                           step_loss)
    assert n == steps == 157
    all_users = np.arange(nu, dtype=np.int32)
    metric = [1, 2, 3, 4, 5]
    cpu_rows = oracle.eval_mf(tr.U, tr.V, all_users, d["train_indptr"], d["train_indices"], d["test_indptr"],
                              d["test_indices"], metric, 20, thread_num=4)
    gpu_rows = ops.eval_mf(dU, dV, dev(all_users), dev(d["train_indptr"]), dev(d["train_indices"]),
                           dev(d["test_indptr"]), dev(d["test_indices"]), metric, 20).cpu().numpy()
    cpu_mean = cpu_rows.astype(np.float64).mean(0).reshape(5, 20)
    gpu_mean = gpu_rows.astype(np.float64).mean(0).reshape(5, 20)
    ndcg10_cpu, ndcg10_gpu = cpu_mean[3, 9], gpu_mean[3, 9]
    assert ndcg10_cpu > 0.02                                   # the epoch really learned something
    assert abs(ndcg10_gpu - ndcg10_cpu) < 1e-5, (ndcg10_gpu, ndcg10_cpu)
    assert np.abs(gpu_mean - cpu_mean).max() < 1e-4           # every metric, every cut-off
