"""GPU parity tests of the tensor-core (tcgen05 / TMEM) evaluator path: bit-identical ranks and
metric rows to the oracle (and therefore to nrc_eval_mf), including masks, ties and candidate
overflow (which must fall back to the exact heap replay)."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from conftest import random_csr

pytestmark = pytest.mark.gpu
ALL = [1, 2, 3, 4, 5]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("K,sw", [(64, 0), (128, 0), (64, 1), (128, 1), (256, 1)])
def test_tcgen05_gemm_building_block(K, sw):
    from neurec_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(K + sw)
    A = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(256, K, device="cuda", generator=g).bfloat16()
    out = torch.zeros(128, 256, device="cuda")
    _lib.check(lib.nrc_tc_gemm_debug(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), K, sw,
                                     ctypes.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T          # bf16 products are exact; only the fp32 accumulation differs
    assert (out.double() - ref).abs().max().item() < 2e-4


def _problem(nu, ni, dim, seed, scale=0.1, int_tables=False):
    rs = np.random.RandomState(seed)
    if int_tables:
        U = rs.randint(-2, 3, size=(nu, dim)).astype(np.float32)
        V = rs.randint(-2, 3, size=(ni, dim)).astype(np.float32)
    else:
        U = (rs.randn(nu, dim) * scale).astype(np.float32)
        V = (rs.randn(ni, dim) * scale).astype(np.float32)
    tp, ti = random_csr(rs, nu, ni, rs.randint(1, 80, nu))
    sp, si = random_csr(rs, nu, ni, rs.randint(1, 12, nu))
    return U, V, tp, ti, sp, si


@pytest.mark.parametrize("nu,ni,dim,K", [(300, 5000, 64, 20), (129, 2049, 128, 10), (500, 12345, 128, 31),
                                         (77, 300, 64, 5)])
def test_tc_eval_bit_exact_vs_oracle(nu, ni, dim, K):
    from neurec_b200 import ops
    U, V, tp, ti, sp, si = _problem(nu, ni, dim, nu + ni)
    users = np.random.RandomState(1).permutation(nu).astype(np.int32)
    tip = np.zeros(nu + 1, np.int64); tip[1:] = np.cumsum(sp[users + 1] - sp[users])
    tix = np.concatenate([si[sp[u]:sp[u + 1]] for u in users])
    want, wranks = oracle.eval_mf(U, V, users, tp, ti, tip, tix, ALL, K, thread_num=4, return_ranks=True)
    args = (dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, K)
    got, ranks = ops.eval_mf_tc(*args, return_ranks=True)
    assert np.array_equal(ranks.cpu().numpy(), wranks)
    assert np.array_equal(got.cpu().numpy(), want)
    got2, ranks2 = ops.eval_mf(*args, return_ranks=True)               # and identical to the SIMT path
    assert torch.equal(ranks, ranks2) and torch.equal(got, got2)


def test_tc_eval_ties_and_overflow_fall_back_to_heap_replay():
    from neurec_b200 import ops
    # integer tables: massive exact ties -> undecidable users -> heap replay must reproduce libstdc++ order
    U, V, tp, ti, sp, si = _problem(140, 900, 64, 5, int_tables=True)
    users = np.arange(140, dtype=np.int32)
    want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, 20, return_ranks=True)
    args = (dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, 20)
    got, ranks = ops.eval_mf_tc(*args, return_ranks=True)
    assert np.array_equal(ranks.cpu().numpy(), wranks) and np.array_equal(got.cpu().numpy(), want)
    # tiny candidate buffer: (almost) every user overflows and is re-done exactly
    U, V, tp, ti, sp, si = _problem(150, 4000, 128, 6)
    users = np.arange(150, dtype=np.int32)
    want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, 20, return_ranks=True)
    got, ranks = ops.eval_mf_tc(dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, 20,
                                return_ranks=True, cand_cap=24)
    assert np.array_equal(ranks.cpu().numpy(), wranks) and np.array_equal(got.cpu().numpy(), want)


def test_tc_eval_candidate_list_heap_replay_matches_reference_heap():
    """Force every user through the candidate-list heap replay (the tie path): the first 2K items seed
    the heap, the candidates are a superset of everything that later enters it, so ranks and metrics
    must still be bit-identical -- also with integer tables where most of the catalogue ties."""
    from neurec_b200 import ops, _lib
    lib = _lib.load()
    try:
        lib.nrc_eval_force_exact(1)
        for (nu, ni, dim, K, ints) in ((200, 6000, 64, 20, False), (130, 3000, 128, 7, False),
                                       (140, 5000, 64, 20, True), (64, 45, 64, 31, False)):
            U, V, tp, ti, sp, si = _problem(nu, ni, dim, 17 + ni, int_tables=ints)
            if ni < 100:
                rs = np.random.RandomState(4)
                tp, ti = random_csr(rs, nu, ni, rs.randint(0, 20, nu))
                sp, si = random_csr(rs, nu, ni, rs.randint(1, 6, nu))
            users = np.arange(nu, dtype=np.int32)
            want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, K, return_ranks=True)
            got, ranks = ops.eval_mf_tc(dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, K,
                                        return_ranks=True)
            assert np.array_equal(ranks.cpu().numpy(), wranks), (nu, ni, dim, K, ints)
            assert np.array_equal(got.cpu().numpy(), want)
            n = ctypes.c_int32(0)
            lib.nrc_eval_last_undecided(ctypes.byref(n))
            assert n.value == nu          # every user went through a heap replay
    finally:
        lib.nrc_eval_force_exact(0)


def test_tc_eval_large_scale_matches_simt_path():
    """200 k items x 2 000 users, d=128: the tensor-core path must agree bit for bit with the SIMT
    fused evaluator (itself pinned on the oracle) -- checks the error-margin argument at scale,
    with score magnitudes that make bf16 rounding errors comparable to the top-K gaps."""
    from neurec_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    nu, ni, dim, K = 2000, 200_000, 128, 20
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
    rs = np.random.RandomState(3)
    tp, ti = random_csr(rs, nu, ni, np.full(nu, 50))
    sp, si = random_csr(rs, nu, ni, np.full(nu, 10))
    users = torch.arange(nu, dtype=torch.int32, device="cuda")
    a = ops.eval_mf_tc(U, V, users, dev(tp), dev(ti), dev(sp), dev(si), ALL, K, return_ranks=True)
    b = ops.eval_mf(U, V, users, dev(tp), dev(ti), dev(sp), dev(si), ALL, K, return_ranks=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])


def test_uni_evaluator_routes_large_catalogues_to_the_tensor_core_path(monkeypatch):
    """UniEvaluator picks nrc_eval_mf_tc above TC_MIN_ITEMS; the printed metric string (the thing
    main.py logs, uni_evaluator.py:150-156) must not change by a character."""
    from neurec_b200.evaluator.uni_evaluator import UniEvaluator
    from neurec_b200 import ops
    nu, ni, dim = 1100, 3000, 64
    U, V, tp, ti, sp, si = _problem(nu, ni, dim, 77)
    train = {u: ti[tp[u]:tp[u + 1]].tolist() for u in range(nu)}
    test = {u: si[sp[u]:sp[u + 1]].tolist() for u in range(nu)}

    class Model:
        def get_eval_tables(self):
            return dev(U), dev(V)
    ev = UniEvaluator(train, test, metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=[5, 10, 20])
    plain = ev.evaluate(Model())
    calls = []
    real = ops.eval_mf_tc
    monkeypatch.setattr(ops, "eval_mf_tc", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    monkeypatch.setattr(ops, "TC_MIN_ITEMS", 1000)
    assert ev.evaluate(Model()) == plain
    assert calls == [1]


@pytest.mark.parametrize("K", [1, 5])
def test_tc_eval_adversarial_bf16_rounding_keeps_the_exact_top1(K):
    """Round-1 verdict's counter-example (tests/test_tc_algorithm_model.py::adversarial_top1_tables):
    bf16 rounding pushes two decoys UP and the exact top-1 DOWN by almost 2^-7 relative each.  With the
    round-1 margin (2^-8) the true top-1 was never a candidate; with eps = 2^-7 + 2^-11 it must be.
    Padded past TC_MIN_ITEMS (16 384) so this is the shape UniEvaluator routes to the tensor cores."""
    from neurec_b200 import ops
    from test_tc_algorithm_model import adversarial_top1_tables
    ni, dim = 16_500, 128
    u, V = adversarial_top1_tables(n_items=ni, d=dim)
    rs = np.random.RandomState(5)
    V[3:] *= rs.uniform(0.2, 1.0, size=(ni - 3, 1)).astype(np.float32)      # distinct fillers, all far below
    place = rs.permutation(ni)                                               # decoys / target anywhere in the stream
    V = np.ascontiguousarray(V[np.argsort(place)])
    target = int(place[2])
    nu = 140
    U = np.stack([u * np.float32(2.0 ** (j % 7 - 3)) for j in range(nu)])   # power-of-two scales keep the roundings
    users = np.arange(nu, dtype=np.int32)
    tp, ti = random_csr(rs, nu, ni, np.full(nu, 3))
    for r in range(nu):                                                      # never mask the three special items
        row = ti[tp[r]:tp[r + 1]]
        assert not (set(row.tolist()) & {int(place[0]), int(place[1]), target})
    sp, si = random_csr(rs, nu, ni, np.full(nu, 4))
    want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, K, return_ranks=True)
    assert (wranks[:, 0] == target).all()
    got, ranks = ops.eval_mf_tc(dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, K,
                                return_ranks=True)
    assert np.array_equal(ranks.cpu().numpy(), wranks)
    assert np.array_equal(got.cpu().numpy(), want)


def test_tc_eval_sixteen_epilogue_warps_give_the_same_bits():
    """The 16-warp epilogue (two threads per user, one per half item tile, own threshold + candidate list
    each) must select exactly what the 8-warp layout selects: ranks and metric rows bit-identical to the
    oracle, also with masks, ties (integer tables) and a catalogue that is not a multiple of the tile."""
    from neurec_b200 import ops
    try:
        ops.eval_tc_epilogue_warps(16)
        for (nu, ni, dim, K, ints) in ((300, 5000, 64, 20, False), (129, 20049, 128, 10, False), (140, 900, 64, 20, True),
                                       (500, 12345, 128, 31, False)):
            U, V, tp, ti, sp, si = _problem(nu, ni, dim, nu + ni + 1, int_tables=ints)
            users = np.arange(nu, dtype=np.int32)
            want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, K, thread_num=4, return_ranks=True)
            got, ranks = ops.eval_mf_tc(dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, K,
                                        return_ranks=True)
            assert np.array_equal(ranks.cpu().numpy(), wranks), (nu, ni, dim, K)
            assert np.array_equal(got.cpu().numpy(), want)
    finally:
        ops.eval_tc_epilogue_warps(8)
