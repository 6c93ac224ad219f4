"""Mash's p-value, pinned independently of the code under test.

Mash: pValue(x, lenRef, lenQuery, kmerSpace, s) = P(X > x-1), X ~ Binomial(s, r),
r = pX pY / (pX + pY - pX pY), pX = 1 / (1 + kmerSpace / len)   (CommandDistance.cpp; GSL gsl_cdf_binomial_Q or
Boost cdf(complement(binomial)) depending on the build -- SURVEY.md App. A.6).  The reference's fixtures only hold
p = 0 and p = 1 (tests/test_correct_workflow.py:22, :35), so the oracle (oracle/mash_oracle.c, Loader's saddle-point
pmf + tail summation) and the CUDA kernel (csrc/ac_dist.cu, the same algorithm) are both checked here against an
exact evaluation in 60-digit arithmetic (mpmath) and against scipy.stats.binom.sf -- north_star's tolerance, 1e-12
relative, written in the asserts.  Vectors: SURVEY.md App. C.4 plus the case the round-1 review found closest to the
bound, (x = 20, n = 50000, r = 5e-4).
"""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import oracle

TOL = 1e-12


def mash_r(k: int, len_ref: int, len_qry: int) -> float:
    space = 4.0 ** k
    px, py = 1.0 / (1.0 + space / len_ref), 1.0 / (1.0 + space / len_qry)
    return px * py / (px + py - px * py)


def exact_Q(k: int, p: float, n: int):
    """P(X > k), X ~ Binomial(n, p), summed over the shorter tail in 60-digit arithmetic (p is the double as given)."""
    with mp.workdps(60):
        p = mp.mpf(p)
        q = 1 - p
        if k >= n:
            return mp.mpf(0)
        if k + 1 > n * p:
            t = mp.binomial(n, k + 1) * p ** (k + 1) * q ** (n - k - 1)
            s, i = t, k + 1
            while i < n:
                t = t * (n - i) / (i + 1) * p / q
                s += t
                i += 1
                if t < s * mp.mpf(10) ** -40:
                    break
            return s
        t = mp.binomial(n, k) * p ** k * q ** (n - k)
        s, i = t, k
        while i > 0:
            t = t * i / (n - i + 1) * q / p
            s += t
            i -= 1
            if t < s * mp.mpf(10) ** -40:
                break
        return 1 - s


def rel_err(got: float, exact) -> float:
    return float(abs((mp.mpf(got) - exact) / exact))


def grid():
    cases = [(20, 5e-4, 50000)]
    r27 = mash_r(27, 12_400_000, 12_400_000)
    assert r27 == 3.4416913751534616e-10
    cases += [(x - 1, r27, 50000) for x in (1, 2, 3, 5, 10)]
    r16 = mash_r(16, 48502, 39937)
    assert r16 == 5.099518672708983e-06
    cases += [(0, r16, 1000), (1, r16, 1000)]
    for n in (100, 1000, 10000, 50000, 1000000):
        for r in (1e-10, 1e-7, 1e-5, 5e-4, 1e-3, 1e-2, 0.1, 0.5):
            m = n * r
            for x in sorted({0, 1, 2, 3, 5, 10, 19, 20, 21, 24, 25, 26, 30, 50, 100, int(m), int(m) + 1, int(1.5 * m) + 1, int(0.5 * m)}):
                if x < n:
                    cases.append((x, r, n))
    return cases


def test_survey_c4_vectors_are_what_an_exact_evaluation_gives():
    r27 = mash_r(27, 12_400_000, 12_400_000)
    for x, want in ((1, 1.7208308814083868e-05), (2, 1.4806083417312364e-10), (3, 8.492643036938254e-16),
                    (5, 1.2572895452470363e-26)):
        assert rel_err(want, exact_Q(x - 1, r27, 50000)) < 1e-13
    r16 = mash_r(16, 48502, 39937)
    for x, want in ((1, 0.005086551137905574), (2, 1.2945554757620216e-05)):
        assert rel_err(want, exact_Q(x - 1, r16, 1000)) < 1e-13


def test_oracle_pvalue_within_1e12_of_exact_and_scipy():
    from scipy.stats import binom
    worst = (0.0, None)
    for k, r, n in grid():
        ex = exact_Q(k, r, n)
        if ex < mp.mpf(10) ** -300:          # denormal / underflow territory: Mash prints 0
            assert oracle.binomial_Q(k, r, n) < 1e-299
            continue
        got = oracle.binomial_Q(k, r, n)
        e = rel_err(got, ex)
        if e > worst[0]:
            worst = (e, (k, r, n, got))
        # scipy (Boost's incomplete beta): an independent double-precision implementation -- itself only good to a few
        # 1e-12 (at (24, 5e-4, 50000) it is 2.5e-12 off the exact value, the oracle 1e-13), so a sanity bound only
        sf = float(binom.sf(k, n, r))
        if sf > 1e-290:
            assert abs(got - sf) <= 1e-11 * sf, (k, r, n, got, sf)
    assert worst[0] < TOL, worst


def test_oracle_dist_pvalue_uses_mash_r():
    """oracle.dist end to end: x common hashes out of s, lengths -> r -> tail."""
    s, x = 1000, 2
    shared = np.arange(1, x + 1, dtype=np.uint64)
    ref = np.concatenate([shared, 10 + 2 * np.arange(s - x, dtype=np.uint64)])
    qry = np.concatenate([shared, 11 + 2 * np.arange(s - x, dtype=np.uint64)])
    d, p, c, n = oracle.dist(ref, qry, s, 16, 48502, 39937)
    assert (c, n) == (x, s)
    assert rel_err(p, exact_Q(x - 1, mash_r(16, 48502, 39937), s)) < TOL
    j = x / s
    assert d == pytest.approx(-math.log(2 * j / (1 + j)) / 16, rel=1e-15)


GPU_CASES = [  # (x common, sketch size, k, len_ref, len_qry)
    (1, 50000, 27, 12_400_000, 12_400_000), (2, 50000, 27, 12_400_000, 12_400_000), (3, 50000, 27, 12_400_000, 12_400_000),
    (5, 50000, 27, 12_400_000, 12_400_000), (10, 50000, 27, 12_400_000, 12_400_000),
    (1, 1000, 16, 48502, 39937), (2, 1000, 16, 48502, 39937),
    # r = 5e-4-ish at n = 50000 (mean 25): both tail branches around the mean, incl. the round-1 review's case x - 1 = 20
    (21, 50000, 16, 4_300_000, 4_300_000), (20, 50000, 16, 4_300_000, 4_300_000), (25, 50000, 16, 4_300_000, 4_300_000),
    (26, 50000, 16, 4_300_000, 4_300_000), (30, 50000, 16, 4_300_000, 4_300_000), (60, 50000, 16, 4_300_000, 4_300_000),
    (5, 50000, 16, 4_300_000, 4_300_000),
    # larger r: hundreds of terms in the tail sums
    (3000, 50000, 16, 500_000_000, 500_000_000), (2700, 50000, 16, 500_000_000, 500_000_000),
    (4000, 50000, 16, 500_000_000, 500_000_000), (40000, 50000, 16, 4_294_967_296, 4_294_967_296),
    (17000, 50000, 16, 4_294_967_296, 4_294_967_296), (16000, 50000, 16, 4_294_967_296, 4_294_967_296),
    (120, 1000, 14, 100_000_000, 50_000_000), (500, 1000, 14, 100_000_000, 50_000_000), (999, 1000, 14, 100_000_000, 50_000_000),
]


@pytest.mark.gpu
def test_gpu_pvalue_within_1e12_of_exact(engine):
    """k5_dist's p-value for pairs built to share exactly x hashes, against the exact tail."""
    from auriclass_b200.msh import MshReference, MshSketch
    worst = (0.0, None)
    for x, s, k, len_ref, len_qry in GPU_CASES:
        shared = np.arange(1, x + 1, dtype=np.uint64)
        ref = np.concatenate([shared, 100_000 + 2 * np.arange(s - x, dtype=np.uint64)])
        qry = np.concatenate([shared, 100_001 + 2 * np.arange(s - x, dtype=np.uint64)])
        m = MshSketch(kmer_size=k, sketch_size=s, hash_bits=64 if k > 16 else 32, references=[MshReference("r", "", len_ref, ref)])
        gr = engine.load_references(m)
        q = engine.sketch_from_hashes(qry, k, s, len_qry)
        row = engine.dist(gr, q)[0]
        gr.free()
        q.free()
        assert (row.common, row.denom) == (x, s)
        ex = exact_Q(x - 1, mash_r(k, len_ref, len_qry), s)
        if ex < mp.mpf(10) ** -300:
            assert row.pvalue < 1e-299
            continue
        e = rel_err(row.pvalue, ex)
        if e > worst[0]:
            worst = (e, (x, s, k, row.pvalue))
        o = oracle.dist(ref, qry, s, k, len_ref, len_qry)
        assert abs(row.pvalue - o[1]) <= TOL * o[1], (x, s, k, row.pvalue, o[1])
    assert worst[0] < TOL, worst
