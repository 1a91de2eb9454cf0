"""Known-answer and distribution tests of the oracle's building blocks (CPU only)."""
import ctypes as C

import numpy as np
import pytest


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert list(orc.philox(ctr, key)) == want


def test_neglog_accuracy(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    ws = rng.integers(0, 2 ** 31, 4000)
    e = np.array([L.orc_neglog_u31(int(w)) for w in ws])
    # chord-table sampler: <= 7.6e-6 above the true value (tools/gen_neglog_table.py), never below
    exact = -np.log((ws + 0.5) / 2.0 ** 31)
    assert np.all(e - exact > -3e-6) and np.all(e - exact < 1.2e-5)
    assert L.orc_neglog_u31(2 ** 31 - 1) >= 0.0
    # monotone non-increasing in the word (needed by nobody, but a cheap sanity check of the table)
    grid = np.array([L.orc_neglog_u31(int(w)) for w in np.linspace(0, 2 ** 31 - 1, 5000).astype(np.int64)])
    assert np.all(np.diff(grid) <= 1e-6)


def test_znorm_matches_inverse_cdf(orc):
    from scipy.special import ndtri
    L = orc.lib()
    rng = np.random.default_rng(1)
    ws = rng.integers(0, 2 ** 32, 20000, dtype=np.uint64)
    z = np.array([L.orc_znorm(int(w)) for w in ws])
    t = ((ws & 0x7FFFFFFF) + 0.5) / 2.0 ** 31
    exact = -ndtri(t / 2.0) * np.where(ws >> 31, -1.0, 1.0)
    # minimax chord table: <= 1.6e-6 (tools/gen_znorm_table.py) + float32 rounding
    assert np.max(np.abs(z - exact)) < 2.5e-6
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.02
    # extreme tails stay finite and ordered; the sign bit mirrors exactly; |z| never negative
    assert 6.0 < L.orc_znorm(0) < 6.6 and -6.6 < L.orc_znorm(0x80000000) < -6.0
    assert L.orc_znorm(0x7FFFFFFF) >= 0.0 and L.orc_znorm(0x7FFFFFFF) < 1e-6
    for w in (1, 12345, 0x40000000, 0x7FFFFFFE):
        assert L.orc_znorm(w) == -L.orc_znorm(w | 0x80000000)
    # every dyadic row of the table, incl. the deepest tails
    for lz in range(32):
        w31 = (1 << (31 - lz)) - 1 if lz < 31 else 0
        exact_t = -ndtri((w31 + 0.5) / 2.0 ** 32)
        assert abs(L.orc_znorm(w31) - exact_t) < 2.5e-6, lz


def test_exp_det(orc):
    L = orc.lib()
    xs = np.linspace(-60, 60, 2001)
    got = np.array([L.orc_exp(float(x)) for x in xs])
    np.testing.assert_allclose(got, np.exp(xs), rtol=4e-16)
    assert L.orc_exp(1000.0) == np.inf and L.orc_exp(-1000.0) == 0.0


def test_threshold_sigmoid_reference_kats(orc):
    """tests/test_synthetic_kw_helpers.py:70-91 pins rust.sigmoid to 4 dp; the thresholded form
    follows src/lib.rs:92-105."""
    L = orc.lib()
    for x, s, t in [(0.0, 1.0, 0.0), (1.0, 1.0, 0.0), (0.5, 3.0, 0.1), (2.0, 25.0, 1.5), (-1.0, 2.0, 0.0)]:
        want = 1.0 / (1.0 + np.exp(-s * (x - t)))
        got = L.orc_threshold_sigmoid(x, 0.0, t, s)  # thresh 0 -> plain sigmoid
        assert abs(got - want) < 1e-15
    th = 0.05
    for x in np.linspace(0, 3, 31):
        r = 1.0 / (1.0 + np.exp(-7.0 * (x - 0.8)))
        h = 2.0 + 1e-10
        tt = min(max(h * th, 0.0), 1.0) / h
        want = min(max((1 + 2 * tt) * r - tt, 0.0), 1.0)
        assert abs(L.orc_threshold_sigmoid(float(x), th, 0.8, 7.0) - want) < 1e-15


def test_sum_array_order_matches_ndarray_unrolled_fold(orc):
    """tests/rust/test_numpy_funcs.py pins sum_array/sum_list values; the summation ORDER follows
    ndarray's 8-way unrolled fold (src/lib.rs:107-111)."""
    L = orc.lib()
    x = np.array([0.1 * i for i in range(1, 20)])
    p = [0.0] * 8
    for i in range(0, 16, 8):
        for j in range(8):
            p[j] += x[i + j]
    want = 0.0
    for a, b in ((0, 4), (1, 5), (2, 6), (3, 7)):
        want += p[a] + p[b]
    for v in x[16:]:
        want += v
    assert L.orc_sum_array(C.c_void_p(x.ctypes.data), C.c_int64(len(x))) == want
    y = np.array([1.0, 2.0, 3.5])
    assert L.orc_sum_array(C.c_void_p(y.ctypes.data), C.c_int64(3)) == 6.5


def test_prob_threshold_equivalence(orc):
    """w <= thr  <=>  w * 2^-32 <= p in float64 (the reference's `rng.random() <= p`)."""
    L = orc.lib()
    rng = np.random.default_rng(3)
    for p in list(rng.random(200)) + [0.0, 1.0, 0.5, 1e-12, 1 - 1e-12]:
        thr = L.orc_prob_threshold(float(p))
        for w in {0, thr, min(thr + 1, 2 ** 32 - 1), max(thr - 1, 0), 2 ** 32 - 1}:
            assert (w <= thr) == (w * 2.0 ** -32 <= p), (p, w, thr)


def test_bid_canonicalisation(orc):
    """round(np.maximum(bid, 0.01), 2) (gymnasium_kw_env.py:215)."""
    L = orc.lib()
    rng = np.random.default_rng(4)
    for b in list(rng.uniform(0, 3, 500)) + [0.0, -1.0, 0.005, 0.015, 0.025, 2.675, 1e7]:
        want = int(np.rint(float(np.round(np.maximum(np.float64(b), 0.01), 2)) * 100))
        assert L.orc_bid_to_cents(float(b)) == want, b


def test_laplace_and_revenue_distributions(orc):
    """Free-running samplers against numpy's own draws of the reference expressions
    (synthetic_kw_helpers.py:66-70,104-113): two-sample KS on 40k draws."""
    from scipy.stats import ks_2samp
    L = orc.lib()
    rng = np.random.default_rng(5)
    loc, scale = 0.62, 0.11
    ws = rng.integers(0, 2 ** 32, 40000, dtype=np.uint64)
    mine = np.array([L.orc_laplace_cents(int(w), loc, scale) for w in ws])
    ref = np.rint(np.around(np.maximum(np.abs(rng.laplace(loc, scale, 40000)), 0.0), 2) * 100)
    assert ks_2samp(mine, ref).pvalue > 1e-3
    mu, sd = 0.97, 0.11
    mine = np.array([L.orc_revenue_cents(int(w), mu, sd) for w in ws])
    ref = np.rint(np.around(np.maximum(rng.normal(mu, sd, 40000), 0.01), 2) * 100)
    assert ks_2samp(mine, ref).pvalue > 1e-3
    vols = np.array([L.orc_volume(int(w), 128.0, 33.0) for w in ws])
    refv = np.floor(np.maximum(rng.normal(128.0, 33.0, 40000), 0.0) + 0.5)
    assert ks_2samp(vols, refv).pvalue > 1e-3
    assert (np.array([L.orc_volume(int(w), 1.0, 5.0) for w in ws[:2000]]) >= 0).all()
