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
