"""CPU model of the tensor-core evaluator's SELECTION logic (tc_eval.cu + eval_tc_finalize_kernel),
independent of CUDA: bf16 rounding emulated in numpy, fp32 accumulation in the worst plausible
orders.  The kernels are tested bit-exactly on the GPU (tests/test_gpu_tc_eval.py); what this file
pins is the mathematics those kernels rely on, on adversarial inputs the GPU tests do not sweep:

  (1) |approximate score - exact score| <= margin/2 with the margin tc_prepare_users_kernel computes;
  (2) main pass (threshold = (K+1)-th best approximate score so far, restarted per item segment):
      every item of the exact top K+1 becomes a candidate;
  (3) finalize pre-filter: candidates within `margin` of the (K+1)-th best approximate candidate
      score still contain the exact top K+1;
  (4) replay pass (threshold = 2K-th best so far): every element that enters the reference's heap
      (evaluate.h:38-41: beats the 2K-th best of the prefix) is a candidate.
"""
import numpy as np
import pytest


def bf16_round(x):
    """round-to-nearest-even of fp32 to bfloat16 (kept as fp32 values), like __float2bfloat16_rn."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32).astype(np.uint64)
    lsb = (b >> 16) & 1
    b = (b + 0x7FFF + lsb) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def exact_scores(u, V):
    """the oracle's definition: fp32 FMA chain over k (emulated: fp64 product+add rounded to fp32)."""
    acc = np.zeros(V.shape[0], np.float32)
    for k in range(V.shape[1]):
        acc = (acc.astype(np.float64) + np.float64(u[k]) * V[:, k].astype(np.float64)).astype(np.float32)
    return acc


def approx_scores(u, V, order):
    """bf16 operands, exact products, fp32 accumulation in a given order ('fwd', 'rev', 'pairs')."""
    ub, Vb = bf16_round(u).astype(np.float64), bf16_round(V).astype(np.float64)
    prods = Vb * ub[None, :]
    if order == "f64":
        return prods.sum(1).astype(np.float32)
    ks = range(V.shape[1]) if order == "fwd" else range(V.shape[1] - 1, -1, -1)
    acc = np.zeros(V.shape[0], np.float32)
    for k in ks:
        acc = (acc.astype(np.float64) + prods[:, k]).astype(np.float32)
        if order == "trunc":   # crude model of a truncating adder: drop one more bit
            acc = (acc.view(np.uint32) & np.uint32(0xFFFFFFFE)).view(np.float32)
    return acc


# bf16 has 8 significant bits: unit roundoff 2^-8 per factor, so two rounded factors move a product by
# up to (2^-7 + 2^-16) of its size; 2^-11 covers the fp32 accumulation (tc_prepare_users_kernel).
EPS_REL = 2.0 ** -7 + 2.0 ** -11
EPS_REL_ROUND1 = 2.0 ** -8 + 2.0 ** -11      # the round-1 constant (assumed a roundoff of 2^-9): NOT rigorous


def margin_of(u, V, eps_rel=EPS_REL):
    vmax = np.sqrt((V.astype(np.float32) ** 2).sum(1, dtype=np.float32)).max()
    return np.float32(2.0) * np.float32(eps_rel) * np.sqrt(np.float32((u * u).sum(dtype=np.float32))) \
        * vmax * np.float32(1.001)


CASES = {
    "gauss": lambda rs, n, d: ((rs.randn(d) * 0.1).astype(np.float32), (rs.randn(n, d) * 0.1).astype(np.float32)),
    "heavy_tail_norms": lambda rs, n, d: ((rs.standard_cauchy(d) * 0.05).astype(np.float32),
                                          (rs.randn(n, d) * np.exp(rs.randn(n, 1) * 1.5)).astype(np.float32)),
    "cancelling": lambda rs, n, d: ((np.tile([1.0, -1.0], d // 2) * (1 + rs.rand(d) * 2 ** -8)).astype(np.float32),
                                    (1.0 + rs.rand(n, d) * 2 ** -7).astype(np.float32)),
    # just below / above a bf16 rounding tie (bf16 spacing in [1,2) is 2^-7, the tie sits at 1 + 2^-8):
    # every factor moves by almost the full unit roundoff 2^-8, all in the same direction
    "worst_rounding": lambda rs, n, d: (np.full(d, 1.0 + 2 ** -8 - 2 ** -20, np.float32),
                                        (np.full((n, d), 1.0 + 2 ** -8 - 2 ** -20) * rs.choice([1, 2, 4], (n, 1))).astype(np.float32)),
    "worst_rounding_up": lambda rs, n, d: (np.full(d, 1.0 + 2 ** -8 + 2 ** -20, np.float32),
                                           (np.full((n, d), 1.0 + 2 ** -8 + 2 ** -20) * rs.choice([1, 2, 4], (n, 1))).astype(np.float32)),
    "tiny_and_huge": lambda rs, n, d: ((rs.randn(d) * 1e-18).astype(np.float32), (rs.randn(n, d) * 1e15).astype(np.float32)),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("order", ["f64", "fwd", "rev", "trunc"])
def test_bf16_error_is_within_half_the_margin(case, order):
    rs = np.random.RandomState(sum(map(ord, case)))        # deterministic per case
    u, V = CASES[case](rs, 400, 128)
    ex, ap = exact_scores(u, V), approx_scores(u, V, order)
    m = margin_of(u, V)
    assert np.all(np.abs(ap.astype(np.float64) - ex.astype(np.float64)) <= 0.5 * float(m)), \
        (case, order, float(np.abs(ap - ex).max()), float(m))


def _stream_candidates(ap, masked, LQ, margin, segments):
    """What tc_candidate_kernel keeps: per contiguous item segment, a running threshold = LQ-th best
    approximate score of the unmasked items seen so far in that segment (-inf until LQ seen)."""
    n = len(ap)
    bounds = np.linspace(0, n, segments + 1).astype(int)
    keep = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        best = []
        for i in range(a, b):
            thr = best[0] if len(best) == LQ else -np.inf        # best is kept ascending, best[0] = LQ-th best
            if masked[i] or not (ap[i] > np.float32(thr) - margin):
                continue
            keep.append(i)
            if ap[i] > thr:
                best.append(ap[i]); best.sort()
                if len(best) > LQ:
                    best.pop(0)
    return np.array(keep, dtype=np.int64)


@pytest.mark.parametrize("case", ["gauss", "heavy_tail_norms", "cancelling", "worst_rounding"])
@pytest.mark.parametrize("segments", [1, 3])
def test_candidate_lists_are_supersets(case, segments):
    rs = np.random.RandomState(7 + segments)
    K, n, d = 5, 1500, 64
    u, V = CASES[case](rs, n, d)
    if case in ("cancelling", "worst_rounding"):
        V = (V * (1 + rs.randint(0, 3, (n, 1)) * 2.0 ** -8)).astype(np.float32)      # many near-ties at bf16 resolution
    masked = rs.rand(n) < 0.05
    ex, ap = exact_scores(u, V), approx_scores(u, V, "fwd")
    exm = np.where(masked, -np.inf, ex)
    m = margin_of(u, V)
    # (2) main pass, LQ = K+1
    cand = _stream_candidates(ap, masked, K + 1, m, segments)
    order = np.argsort(-exm, kind="stable")
    cut = exm[order[K]]                                            # exact (K+1)-th best value
    must = np.where(exm >= cut)[0] if np.isfinite(cut) else np.where(np.isfinite(exm))[0]
    assert set(order[:K + 1]) <= set(cand) and set(must) <= set(cand)          # ties at the cut included
    # (3) finalize pre-filter on the candidates' approximate scores
    ac = np.sort(ap[cand])[::-1]
    a_cut = (ac[K] if len(ac) > K else -np.inf) - m
    survivors = cand[~(ap[cand] < a_cut)]
    assert set(order[:K + 1]) <= set(survivors)
    assert len(survivors) <= len(cand)
    # (4) replay pass, LQ = 2K: everything that enters the reference's heap is a candidate
    L = 2 * K
    cand2 = set(_stream_candidates(ap, masked, L, m, segments).tolist())
    heap = sorted(exm[:L].tolist())                                # values only: entering depends on the root value
    for i in range(L, n):
        if exm[i] > heap[0]:                                       # evaluate.h:40-41, strict >
            assert i in cand2, (case, i)
            heap[0] = exm[i]; heap.sort()


def adversarial_top1_tables(n_items=64, d=128):
    """Round-1 verdict's counter-example: two decoys whose bf16 scores round UP by almost 2^-7 relative,
    and the true top-1 (item 2) whose bf16 score rounds DOWN by as much.  K = 1."""
    lo, hi = np.float32(1 + 2.0 ** -8 - 2.0 ** -16), np.float32(1 + 2.0 ** -8 + 2.0 ** -16)
    assert d >= 128
    u = np.zeros(d, np.float32); u[:63] = lo; u[63:126] = hi; u[126] = 0.5
    A = np.zeros(d, np.float32); A[:63] = lo; A[126] = 0.25      # lo*lo: both factors round down
    B = np.zeros(d, np.float32); B[63:126] = hi                  # hi*hi: both factors round up
    V = np.zeros((n_items, d), np.float32)
    V[0], V[1], V[2] = B, B, A
    V[3:] = A * np.float32(0.5)                                  # filler far below
    return u, V


def test_round1_margin_drops_the_exact_top1_and_the_rigorous_one_keeps_it():
    u, V = adversarial_top1_tables()
    ex, ap = exact_scores(u, V), approx_scores(u, V, "fwd")
    assert np.argmax(ex) == 2 and ex[2] > ex[0] == ex[1]          # item 2 is the exact top-1 ...
    assert ap[0] > ap[2] and ap[1] > ap[2]                        # ... but bf16 ranks both decoys above it
    masked = np.zeros(len(ex), bool)
    K = 1
    old = _stream_candidates(ap, masked, K + 1, margin_of(u, V, EPS_REL_ROUND1), 1)
    assert 2 not in set(old.tolist())                             # the round-1 constant loses the true top-1
    new = _stream_candidates(ap, masked, K + 1, margin_of(u, V), 1)
    assert {0, 1, 2} <= set(new.tolist())
    # (1) holds with the rigorous constant and is violated by the old one on this input
    err = np.abs(ap.astype(np.float64) - ex.astype(np.float64)).max()
    assert err <= 0.5 * float(margin_of(u, V)) and err > 0.5 * float(margin_of(u, V, EPS_REL_ROUND1))


def test_single_factor_worst_case_exceeds_the_round1_bound():
    u = np.full(128, 1 + 2.0 ** -8 - 2.0 ** -20, np.float32)
    V = u[None, :].copy()
    ex, ap = exact_scores(u, V), approx_scores(u, V, "f64")
    err = float(abs(np.float64(ap[0]) - np.float64(ex[0])))
    assert err > 0.5 * float(margin_of(u, V, EPS_REL_ROUND1))
    assert err <= 0.5 * float(margin_of(u, V))
