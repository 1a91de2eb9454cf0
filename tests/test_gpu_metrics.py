"""adc_ideal_profit (the ideal-profit estimator behind AKNCP / NCP) against the reference.

* tests/golden/ideal_profit.npz holds what the unmodified reference's
  ``get_implicit_kw_bid_cpc_impressions`` / ``get_max_expected_bid_profits``
  (experiment_metrics.py:20-61) returned for a drifted keyword set, together with the 2048
  competitor bids it sampled per keyword: fed the same samples, the kernel must return the same
  impression rates (exact), prices and profits (1e-12 relative: the reference adds float64 dollars,
  the kernel exact integer cents), positive shares and arg-max bids.
* random cases against a numpy restatement of the same lines (sort / searchsorted / cumsum).
* free-running (Philox samples): rates against the closed-form folded-Laplace CDF.
"""
import os

import numpy as np
import pytest

import golden_io

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _env(cols, E=1, per_env=False, **kw):
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    table = kwm.KeywordTable(kwm.IMPLICIT, *[np.asarray(cols[n], np.float64) for n in golden_io.PARAMS])
    env = VectorBiddingSimulation(E, num_keywords=table.K, keywords=table, device="cuda", seed=3, budget=1e5,
                                  shared_keywords=not per_env, **kw)
    env.reset()
    return env


def _numpy_estimator(samples, grid, vol, bctr, sctr, mrev):
    """experiment_metrics.py:27-35, 51-60 for one keyword."""
    second = np.sort(samples)
    n = len(second)
    idx = np.searchsorted(second, grid, side="right")
    rate = idx / n
    cpc = (np.cumsum(second) / np.arange(1, n + 1, 1))[np.minimum(idx, n - 1)]
    prof = np.maximum(vol * rate * bctr * (sctr * mrev - cpc), 0.0)
    return rate, cpc, max([0.0, prof.max()]), np.sum(prof > 0) / len(cpc), int(np.argmax(prof))


def test_reference_golden_same_samples():
    from adcraft_b200 import metrics as m
    z = np.load(os.path.join(HERE, "golden", "ideal_profit.npz"))
    env = _env({n: z["kw_" + n] for n in golden_io.PARAMS})
    samples = torch.from_numpy(z["samples_cents"][None]).cuda().contiguous()
    out = m.ideal_profit(env, z["allowed_bids"], samples.shape[-1], samples_cents=samples, profile=True)
    assert np.array_equal(out["impression_rate"][0].cpu().numpy(), z["impression_rate"])
    np.testing.assert_allclose(out["expected_cpc"][0].cpu().numpy(), z["expected_cpc"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(out["ideal"][0].cpu().numpy(), z["ideal_profit"], rtol=1e-11, atol=1e-12)
    assert np.array_equal(out["positive_frac"][0].cpu().numpy(), z["positive_frac"])
    assert np.array_equal(out["best_bid_index"][0].cpu().numpy(), z["best_bid_index"])
    assert (z["ideal_profit"] > 0).any() and (z["ideal_profit"] == 0).any()  # both branches of metrics.py:58


@pytest.mark.parametrize("n_samples,E", [(2048, 3), (100, 2), (1, 1), (4097, 2)])
def test_random_cases_against_numpy_restatement(n_samples, E):
    from adcraft_b200 import metrics as m
    from conftest import make_implicit_table
    rng = np.random.default_rng(n_samples)
    K = 11
    table = make_implicit_table(rng, K, 64, E=E)
    env = _env({n: getattr(table, n) for n in golden_io.PARAMS}, E=E, per_env=True)
    # competitor bids incl. values far above the exact range of the counting sort and duplicates
    samples = np.rint(np.abs(rng.laplace(0.6, 0.5, (E, K, n_samples))) * 100).astype(np.int32)
    samples[:, 0] = 0
    samples[:, 1] = 10 ** 6
    samples[:, 2, ::2] = 700
    grid = np.concatenate([[0.0], np.arange(0.01, 3.00, 0.01), [4.99, 5.05]])
    out = m.ideal_profit(env, grid, n_samples, samples_cents=torch.from_numpy(samples).cuda(), profile=True)
    for e in range(E):
        for k in range(K):
            rate, cpc, best, frac, arg = _numpy_estimator(
                samples[e, k] / 100.0, grid, table.vol_mean[e, k], table.ctr[e, k], table.cvr[e, k], table.rev_mean[e, k])
            assert np.array_equal(out["impression_rate"][e, k].cpu().numpy(), rate), (e, k)
            np.testing.assert_allclose(out["expected_cpc"][e, k].cpu().numpy(), cpc, rtol=1e-11, atol=1e-15)
            assert abs(float(out["ideal"][e, k]) - best) <= 1e-10 * max(1.0, best)
            assert abs(float(out["positive_frac"][e, k]) - frac) < 1e-15
            got = int(out["best_bid_index"][e, k])
            if got != arg:  # a tie broken by the float rounding of the two price sums
                p = np.maximum(table.vol_mean[e, k] * rate * table.ctr[e, k] * (table.cvr[e, k] * table.rev_mean[e, k] - cpc), 0)
                assert abs(p[got] - p[arg]) <= 1e-10 * max(1.0, p[arg])


def test_free_running_rates_follow_the_model():
    """Philox-drawn samples: the impression rate of bid b estimates P(round(|Laplace|, 2) <= b)."""
    from adcraft_b200 import metrics as m
    loc, scale = np.array([0.5, 0.9, 0.3]), np.array([0.05, 0.2, 0.1])
    K = 3
    cols = dict(vol_mean=np.full(K, 128.0), vol_std=np.full(K, 8.0), p1=loc, p2=scale, ctr=np.full(K, 0.5),
                cvr=np.full(K, 0.8), rev_mean=np.full(K, 1.0), rev_std=np.full(K, 0.2))
    env = _env(cols)
    n = 65536
    out = m.ideal_profit(env, n_samples=n, profile=True)
    grid = m.DEFAULT_BID_GRID
    tau = (np.floor(grid * 100 + 1e-9) + 0.5) / 100.0  # cents <= c  <=>  |x| < (c + 0.5) / 100
    cdf = lambda x, l, s: np.where(x < l, 0.5 * np.exp((x - l) / s), 1 - 0.5 * np.exp(-(x - l) / s))
    for k in range(K):
        p = cdf(tau, loc[k], scale[k]) - cdf(-tau, loc[k], scale[k])
        got = out["impression_rate"][0, k].cpu().numpy()
        assert np.max(np.abs(got - p)) < 4.5 * 0.5 / np.sqrt(n) + 2e-5
    two = m.ideal_profit(env, n_samples=2048, step=7)["ideal"]
    again = m.ideal_profit(env, n_samples=2048, step=7)["ideal"]
    assert torch.equal(two, again) and float(two.max()) > 0


def test_grid_beyond_the_exact_range_is_refused():
    from adcraft_b200 import _capi, metrics as m
    cols = dict(vol_mean=[1.0], vol_std=[1.0], p1=[0.5], p2=[0.1], ctr=[0.5], cvr=[0.5], rev_mean=[1.0], rev_std=[0.1])
    env = _env(cols)
    with pytest.raises(_capi.AdcError, match="grid bid"):
        m.ideal_profit(env, np.array([0.5, 5.2]))


@pytest.mark.gpu
@pytest.mark.parametrize("K,E,per_env", [(100, 300, False), (7, 33, True), (64, 129, False), (1, 5, False), (501, 40, True)])
def test_episode_metrics_kernel_matches_torch_form(K, E, per_env):
    """adc_episode_metrics (a warp per env: AKNCP = np.median over the keywords of mean profit / mean
    ideal, NCP = sum / sum; experiment_metrics.py:64-83) against the eager torch form of the same
    window -- odd and even K (np.median averages the two middle values), ties, ideal profits <= 0,
    shared and per-env ideal tables; the accumulators are zeroed for the next window."""
    from adcraft_b200 import metrics as M
    from adcraft_b200.vector_env import VectorBiddingSimulation
    from conftest import make_implicit_table
    rng = np.random.default_rng(K + E)
    table = make_implicit_table(rng, K, 40)
    mk = lambda: VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e6, device="cuda", seed=9,
                                         episode_profit=True)
    a, b = mk(), mk()
    a.reset(); b.reset()
    steps = 4
    bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2)).cuda()
    for _ in range(steps):
        a.step({"keyword_bids": bids}); b.step({"keyword_bids": bids})
    ideal = torch.from_numpy(rng.uniform(-0.5, 3.0, (E if per_env else 1, K))).cuda()
    ideal[:, ::5] = 0.0                                    # ideal <= 0 counts as 1
    if K >= 4:                                             # ties around the middle
        for env in (a, b):
            env._out["episode_profit_cents"][:, 1] = env._out["episode_profit_cents"][:, 2]
        ideal[:, 1] = ideal[:, 2]
    assert int(a._out["episode_profit_cents"].abs().sum()) > 0
    va = M.episode_summary_vector(a, steps, ideal, use_kernel=True)
    vb = M.episode_summary_vector(b, steps, ideal, use_kernel=False)
    np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=1e-11, atol=1e-9)
    assert float(va[5]) == E
    assert int(a._out["episode_profit_cents"].abs().sum()) == 0 and int(b._out["episode_profit_cents"].abs().sum()) == 0
