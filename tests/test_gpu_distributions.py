"""Free-running Philox mode against the reference's *distributions* (north_star: KS / chi-square on
per-keyword outcome distributions).  The reference side is numpy evaluating the reference's own
expressions (synthetic_kw_helpers.py:66-77,104-113; src/lib.rs:314-325; bidding_simulation.py:86-117)
with numpy's Generator -- the same calls the reference makes -- for one keyword at a time; the
budget is large so sub-steps do not matter.  Fixed seeds; Bonferroni-corrected thresholds."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _reference_day_samples(rng, n, vol_mean, vol_std, loc, scale, ctr, cvr, rev_mean, rev_std, bid):
    out = np.zeros((n, 5))
    for i in range(n):
        raw = rng.normal(vol_mean, vol_std)
        v = int(np.floor(max(raw, 0.0) + 0.5))
        comp = np.around(np.maximum(np.abs(rng.laplace(loc, scale, (1, v))), 0.0).astype(float), 2).ravel()
        costs = comp[bid > comp]
        clicked = rng.random(len(costs)) <= ctr
        paid = costs[clicked]
        conv = rng.random(len(paid)) <= cvr
        revs = np.around(np.maximum(rng.normal(rev_mean, rev_std, int(conv.sum())), 0.01).astype(float), 2)
        out[i] = (len(costs), len(paid), conv.sum(), paid.sum(), revs.sum())
    return out


def test_outcome_distributions_match_reference_expressions():
    from scipy.stats import ks_2samp
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(2025)
    K, E, steps = 6, 4096, 4
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 64, "conversion_rate": 0.5})
    bids = np.array([0.45, 0.6, 0.75, 0.9, 1.2, 0.3])
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda", seed=777,
                                  obs_dtype=torch.float64)
    env.reset()
    got = []
    tb = torch.from_numpy(np.tile(bids, (E, 1))).cuda()
    for _ in range(steps):
        obs = env.step({"keyword_bids": tb})[0]
        got.append(np.stack([obs[k].cpu().numpy().astype(np.float64) for k in
                             ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue")], -1))
    got = np.concatenate(got)  # [steps*E, K, 5]
    n_tests = K * 5
    alpha = 1e-3 / n_tests
    for k in range(K):
        ref = _reference_day_samples(rng, 6000, table.vol_mean[k], table.vol_std[k], table.p1[k], table.p2[k],
                                     table.ctr[k], table.cvr[k], table.rev_mean[k], table.rev_std[k], bids[k])
        for j, name in enumerate(("impressions", "clicks", "conversions", "cost", "revenue")):
            p = ks_2samp(got[:, k, j], ref[:, j]).pvalue
            assert p > alpha, f"keyword {k} {name}: KS p={p:.2e}"
        # first moments within 4 standard errors
        for j in range(5):
            a, b = got[:, k, j], ref[:, j]
            se = np.sqrt(a.var() / len(a) + b.var() / len(b)) + 1e-12
            assert abs(a.mean() - b.mean()) < 4.5 * se, (k, j, a.mean(), b.mean())


def test_win_click_conversion_rates_chi_square():
    """Per-auction event frequencies against the closed-form probabilities of the reference's
    model: P(win) = P(round(|Laplace|,2) < bid), P(click|win) = ctr, P(conv|click) = cvr."""
    from scipy.stats import chisquare
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    K, E = 4, 8192
    loc = np.array([0.5, 0.7, 0.9, 0.4]); scale = np.array([0.05, 0.12, 0.2, 0.1])
    table = kwm.KeywordTable(kwm.IMPLICIT, np.full(K, 100.0), np.full(K, 1e-9), loc, scale,
                             np.array([0.2, 0.5, 0.8, 0.35]), np.array([0.1, 0.5, 0.9, 0.65]),
                             np.full(K, 1.0), np.full(K, 0.2))
    bids = np.array([0.52, 0.7, 0.61, 0.55])
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda", seed=31337)
    env.reset()
    obs = env.step({"keyword_bids": torch.from_numpy(np.tile(bids, (E, 1))).cuda()})[0]
    I = obs["impressions"].sum(0).cpu().numpy().astype(float)
    B = obs["buyside_clicks"].sum(0).cpu().numpy().astype(float)
    S = obs["sellside_conversions"].sum(0).cpu().numpy().astype(float)
    N = 100.0 * E
    # win iff round(|x|*100) < bid_cents  <=>  |x| < (bid_cents - 0.5)/100
    tau = (np.rint(bids * 100) - 0.5) / 100.0
    cdf = lambda x: np.where(x < loc, 0.5 * np.exp((x - loc) / scale), 1 - 0.5 * np.exp(-(x - loc) / scale))
    p_win = cdf(tau) - cdf(-tau)
    for k in range(K):
        for obs_n, tot, p in ((I[k], N, p_win[k]), (B[k], I[k], table.ctr[k]), (S[k], B[k], table.cvr[k])):
            stat = chisquare([obs_n, tot - obs_n], [tot * p, tot * (1 - p)])
            assert stat.pvalue > 1e-4 / (3 * K), (k, obs_n, tot, p, stat.pvalue)


def _threshold_sigmoid(bid, thresh, intercept, slope):
    """src/lib.rs:92-105, 290-300."""
    halver = 2.0 + 1e-10
    t = min(max(halver * thresh, 0.0), 1.0) / halver
    r = 1.0 / (1.0 + np.exp(-slope * (bid - intercept)))
    return min(max((1.0 + 2.0 * t) * r - t, 0.0), 1.0)


def _reference_explicit_day_samples(rng, n, vol_mean, vol_std, intercept, slope, ctr, cvr, rev_mean, rev_std, bid,
                                    thresh=0.05):
    """One ExplicitKeyword day in the reference's own expressions (classes:493-538, lib.rs:53-76,
    bsim:44-120,151-167): 24 lanes, Binomial impressions, cost_create, the phantom zero-cost slot."""
    p = _threshold_sigmoid(bid, thresh, intercept, slope)
    xs = np.sqrt(bid)
    out = np.zeros((n, 5))
    for i in range(n):
        v = int(np.floor(max(rng.normal(vol_mean, vol_std), 0.0) + 0.5))
        q = v // 24
        I = B = S = 0
        cost = rev = 0.0
        for t in range(24):
            nt = v - 23 * q if t == 0 else q
            imp = rng.binomial(nt, p)
            costs = (np.clip(xs / 4.0 + 2.2 + rng.normal(0.0, 1e-10 + xs / 6.0, imp), 0.0, 4.4)
                     if imp >= 1 else np.zeros(1))  # classes:514-515: one phantom slot
            clicked = rng.random(len(costs)) <= ctr
            conv = rng.random(int(clicked.sum())) <= cvr
            revs = np.around(np.maximum(rng.normal(rev_mean, rev_std, int(conv.sum())), 0.01), 2)
            I += imp; B += int(clicked.sum()); S += int(conv.sum())
            cost += costs[clicked].sum(); rev += revs.sum()
        out[i] = (I, B, S, cost, rev)
    return out


def test_explicit_outcome_distributions_match_reference_expressions():
    """Explicit keywords in free-running mode: Bernoulli-sum impressions, chord-table normal costs,
    phantom slots -- against the reference's expressions evaluated by numpy."""
    from scipy.stats import ks_2samp
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(515)
    K, E, steps = 5, 4096, 3
    table = kwm.sample_random_keywords(K, rng)
    bids = np.array([0.4, 0.9, 1.5, 2.2, 2.9])
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda", seed=4321,
                                  obs_dtype=torch.float64)
    env.reset()
    tb = torch.from_numpy(np.tile(bids, (E, 1))).cuda()
    got = []
    for _ in range(steps):
        obs = env.step({"keyword_bids": tb})[0]
        got.append(np.stack([obs[k].cpu().numpy().astype(np.float64) for k in
                             ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue")], -1))
    got = np.concatenate(got)
    alpha = 1e-3 / (K * 5)
    for k in range(K):
        ref = _reference_explicit_day_samples(rng, 5000, table.vol_mean[k], table.vol_std[k], table.p1[k], table.p2[k],
                                              table.ctr[k], table.cvr[k], table.rev_mean[k], table.rev_std[k], bids[k])
        for j, name in enumerate(("impressions", "clicks", "conversions", "cost", "revenue")):
            p = ks_2samp(got[:, k, j], ref[:, j]).pvalue
            assert p > alpha, f"explicit keyword {k} {name}: KS p={p:.2e}"
            a, b = got[:, k, j], ref[:, j]
            se = np.sqrt(a.var() / len(a) + b.var() / len(b)) + 1e-12
            assert abs(a.mean() - b.mean()) < 4.5 * se, (k, name, a.mean(), b.mean())
    assert (got[:, :, 1] > got[:, :, 0]).any()  # phantom slots: clicks can exceed impressions


def test_explicit_click_cost_distribution():
    """rust.cost_create (src/lib.rs:53-67) per impression: with a volume of one auction per day the
    day's cost, when positive, IS one click's cost -- one-sample KS against the clamped normal
    clamp(sqrt(b)/4 + 2.2 + N(0, 1e-10 + sqrt(b)/6), 0, 4.4); impressions ~ Bernoulli(thresholded
    sigmoid) by chi-square."""
    from scipy.stats import chisquare, kstest, norm
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    K, E = 4, 16384
    bids = np.array([0.25, 1.0, 2.0, 2.95])
    intercept, slope = np.array([0.3, 0.9, 2.1, 2.9]), np.array([9.0, 4.0, 4.0, 6.0])
    table = kwm.KeywordTable(kwm.EXPLICIT, np.full(K, 1.0), np.full(K, 1e-9), intercept, slope,
                             np.full(K, 1.0), np.full(K, 0.5), np.full(K, 1.0), np.full(K, 0.2))
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda", seed=99,
                                  obs_dtype=torch.float64)
    env.reset()
    tb = torch.from_numpy(np.tile(bids, (E, 1))).cuda()
    imps, costs = [], []
    for _ in range(3):
        obs = env.step({"keyword_bids": tb})[0]
        imps.append(obs["impressions"].cpu().numpy()); costs.append(obs["cost"].cpu().numpy())
    imps, costs = np.concatenate(imps), np.concatenate(costs)
    for k in range(K):
        p = _threshold_sigmoid(bids[k], 0.05, intercept[k], slope[k])
        n = len(imps)
        got = float(imps[:, k].sum())
        assert 0.02 < p < 0.98, (k, p)
        assert chisquare([got, n - got], [n * p, n * (1 - p)]).pvalue > 1e-4 / K, (k, got / n, p)
        c = costs[imps[:, k] == 1, k]  # ctr = 1: every impression is clicked, the phantom slots cost 0
        xs = np.sqrt(bids[k])
        mu, sd = xs / 4.0 + 2.2, 1e-10 + xs / 6.0
        inner = c[(c > 0.0) & (c < 4.4)]  # the clamp puts atoms at 0 and 4.4
        lo, hi = norm.cdf(0.0, mu, sd), norm.cdf(4.4, mu, sd)
        assert kstest(inner, lambda x: (norm.cdf(x, mu, sd) - lo) / (hi - lo)).pvalue > 1e-4 / K, k
        assert abs((c >= 4.4).mean() - (1 - hi)) < 4.5 * np.sqrt(max((1 - hi) * hi, 1e-9) / len(c)) + 1e-4
