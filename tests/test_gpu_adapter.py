"""The E = 1 adapter, flat wrapper and device metrics on the GPU (reference API shape)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_single_env_adapter_api_and_dtypes():
    """Mirrors adcraft/tests/test_env.py: reset/step run, obs are in the observation space after the
    reference's own dtype cast (test_env.py:60-69), reward is a float, flags are bools."""
    from adcraft_b200.gymnasium_kw_env import BiddingSimulation, bidding_sim_creator
    env = BiddingSimulation()
    reset_obs, info = env.reset(seed=1)
    assert env.observation_space.contains(reset_obs) and "keyword_params" in info
    action = env.action_space.sample()
    obs, reward, terminated, truncated, info = env.step(action)
    assert isinstance(reward, float) and isinstance(terminated, bool) and isinstance(truncated, bool)
    assert set(info) == {"bids", "bidding_outcomes", "keyword_params"}
    assert obs["impressions"].dtype == np.int64 and obs["cost"].dtype == np.float64
    assert obs["cumulative_profit"].shape == (1,) and obs["days_passed"].tolist() == [1]
    cast = {k: obs[k].astype(v.dtype) for k, v in reset_obs.items()}
    assert env.observation_space.contains(cast)
    assert abs(obs["cumulative_profit"][0] - reward) < 1e-9
    env2 = bidding_sim_creator(dict(keyword_config={"mean_volume": 16, "conversion_rate": 0.5}, num_keywords=2,
                                    max_days=3))
    _, info2 = env2.reset(seed=0)
    assert "imp_intercept: 0.6459721981904619" in info2["keyword_params"]  # notebook golden
    done = False
    for _ in range(3):
        o, r, done, tr, _ = env2.step({"keyword_bids": np.array([0.75, 0.75]), "budget": 100000})
    assert done and o["days_passed"][0] == 3


def test_adapter_budget_aliasing_follows_action_type():
    from adcraft_b200.gymnasium_kw_env import BiddingSimulation
    cfg = dict(keyword_config={"mean_volume": 128, "conversion_rate": 0.8}, num_keywords=20, seed=3)
    a, b = BiddingSimulation(**cfg), BiddingSimulation(**cfg)
    a.reset(seed=7); b.reset(seed=7)
    bids = np.full(20, 0.9)
    oa = a.step({"keyword_bids": bids, "budget": 40.0})[0]
    ob = b.step({"keyword_bids": bids, "budget": np.array([40.0])})[0]
    # the ndarray budget is charged twice per click: about half the spend before the day ends
    assert oa["cost"].sum() <= 40.0 + 1e-9 and ob["cost"].sum() <= 20.0 + 1.5
    assert ob["cost"].sum() < oa["cost"].sum()
    assert isinstance(b.budget, np.ndarray) and float(b.budget[0]) <= 0.0 + 1.5


def test_flat_wrapper_round_trip():
    from adcraft_b200.gymnasium_kw_env import BiddingSimulation
    from adcraft_b200.wrappers import FlatArrayWrapper, observation_slices
    env = FlatArrayWrapper(BiddingSimulation(num_keywords=5))
    flat, _ = env.reset(seed=2)
    assert flat.shape == (27,)
    act = np.concatenate([[1000.0], np.full(5, 1.2)]).astype(np.float32)
    flat, reward, term, trunc, info = env.step(act)
    sl = observation_slices(5)
    assert flat.shape == (27,) and flat[sl["days_passed"]][0] == 1
    assert abs(flat[sl["revenue"]].sum() - flat[sl["cost"]].sum() - reward) < 1e-9


def test_device_metrics_against_host_definitions():
    from adcraft_b200 import keywords as kwm, metrics as m
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(5)
    K, E = 12, 16
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 64, "conversion_rate": 0.8})
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e5, device="cuda", seed=1,
                                  obs_dtype=torch.float64)
    env.reset()
    ideal = m.ideal_profit(env)["ideal"][0]
    acc = m.MetricAccumulator(E, K, "cuda")
    profits, bids = [], torch.full((E, K), 0.75, dtype=torch.float64, device="cuda")
    for t in range(10):
        obs, reward, term, trunc, _ = env.step({"keyword_bids": bids})
        acc.update(obs, reward, ideal=ideal[None], done=term)
        profits.append((obs["revenue"] - obs["cost"]).cpu().numpy())
    pe = acc.per_env()
    P, I = np.stack(profits), np.tile(ideal.cpu().numpy(), (10, 1))
    for e in range(E):
        assert abs(float(pe["akncp"][e]) - m.compute_AKNCP(P[:, e], I)) < 1e-9
        assert abs(float(pe["ncp"][e]) - m.compute_NCP(P[:, e], I)) < 1e-9
    s = m.summarize(m.reduce_metrics(acc.summary_vector()))
    assert s["n_envs"] == E


@pytest.mark.parametrize("K,E,budget", [(33, 70, 60.0), (120, 9, 150.0), (20, 40, 1e6)])
def test_step_host_zero_copy_equals_staged_copy(K, E, budget):
    """The fused (UVA zero-copy) host round trip returns exactly what the staged one returns --
    with binding budgets too, where the exact serial walk keeps its running counts in device
    scratch (adc_scratch.acc_*) instead of the host-mapped outputs; K = 120 is beyond the serial
    kernel's shared-memory keyword cache."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(8)
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 64, "conversion_rate": 0.8})
    mk = lambda: VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=budget, device="cuda", seed=4,
                                         max_days=2)
    a, b = mk(), mk()
    a.reset(); b.reset()
    for step in range(3):
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2).astype(np.float32)).pin_memory()
        ha = {k: v.clone() for k, v in a.step_host(bids, zero_copy=True).items()}
        hb = b.step_host(bids, zero_copy=False)
        for k in hb:
            assert torch.equal(ha[k], hb[k]), (k, step)
        assert ha["impressions"].sum() > 0


def test_rollout_example_runs():
    """examples/rollout_collect.py: MLP policy on flat observations, metric accumulation, reduce."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "rollout_collect.py"), "--envs", "256",
                          "--keywords", "20", "--steps", "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["n_envs"] == 256 and d["episodes"] == 256 and d["units_per_s"] > 0


def test_bidding_outcomes_detail_is_consistent():
    """info["bidding_outcomes"] (rust.repr_outcomes_py format): per-click costs, per-click revenues,
    impression share -- consistent with the observation of the same step."""
    import ast
    from adcraft_b200.gymnasium_kw_env import BiddingSimulation
    for env in (BiddingSimulation(num_keywords=6),
                BiddingSimulation(keyword_config={"mean_volume": 64, "conversion_rate": 0.5}, num_keywords=6)):
        env.reset(seed=4)
        obs, reward, _, _, info = env.step({"keyword_bids": np.full(6, 0.9), "budget": 500.0})
        outcomes = ast.literal_eval(info["bidding_outcomes"])
        assert len(outcomes) == 6
        for k, o in enumerate(outcomes):
            assert o["impressions"] == obs["impressions"][k] and o["buyside_clicks"] == obs["buyside_clicks"][k]
            assert len(o["costs"]) == o["buyside_clicks"] == len(o["revenues_per_cost"])
            assert len(o["revenues"]) == o["sellside_conversions"] == obs["sellside_conversions"][k]
            assert abs(sum(o["costs"]) - obs["cost"][k]) < 1e-9 and abs(sum(o["revenues"]) - obs["revenue"][k]) < 1e-9
            assert [r for r in o["revenues_per_cost"] if r > 0] == o["revenues"]
            assert 0.0 <= o["impression_share"] <= 1.0 + 1e-12
            assert abs(o["profit"] - (obs["revenue"][k] - obs["cost"][k])) < 1e-9
        assert abs(sum(o["profit"] for o in outcomes) - reward) < 1e-9


def test_device_side_keyword_sampling_matches_host_distributions():
    """per-env keyword sets drawn on the GPU follow the same distributions as the host
    factory that reproduces the reference's draws (two-sample KS per parameter)."""
    from scipy.stats import ks_2samp
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    cfg = {"mean_volume": 64, "conversion_rate": 0.3}
    E, K = 64, 200
    g = torch.Generator(device="cuda").manual_seed(11)
    cols = kwm.sample_implicit_keywords_device(E, K, cfg, torch.device("cuda", 0), generator=g)
    host = kwm.sample_implicit_keywords_from_quantiles(E * K // 4, np.random.default_rng(3), cfg)
    for n in kwm.PARAM_NAMES:
        a, b = cols[n].cpu().numpy().ravel(), getattr(host, n).ravel()
        if np.unique(b).size == 1:
            assert np.all(a == b[0]), n
        else:
            assert ks_2samp(a, b).pvalue > 1e-4, n
    env = VectorBiddingSimulation(E, num_keywords=K, device="cuda", seed=1, budget=1e6)
    env.install_device_keywords(cols)
    obs = env.step({"keyword_bids": torch.full((E, K), 0.8, device="cuda")})[0]
    assert int(obs["impressions"].sum()) > 0


@pytest.mark.parametrize("K,E,budget,chunks,drift", [(33, 70, 60.0, 4, False), (100, 256, 1e5, 4, True),
                                                     (20, 40, 1e6, 3, False), (7, 5, 3.0, 8, True)])
def test_step_host_pipelined_equals_device_step(K, E, budget, chunks, drift):
    """adc_step_host (chunked H2D -> kernels -> row packing -> D2H on one stream per chunk) returns, in
    its compact host rows, exactly what the plain device step computes -- budgets that bind (the
    exact serial walk runs per chunk on its share of the workspace), drift, ragged chunk sizes."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(K)
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 64, "conversion_rate": 0.8})
    mk = lambda: VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=budget, device="cuda", seed=4,
                                         max_days=3, updater_mask=[True] * K if drift else None)
    a, b, c, d = mk(), mk(), mk(), mk()
    a.reset(); b.reset(); c.reset(); d.reset()
    for step in range(5):
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2).astype(np.float32)).pin_memory()
        h = a.step_host_pipelined(bids, n_chunks=chunks)
        # the fused form: rows packed by the kernels straight into pinned host memory (adc_step_out.rows)
        hr = c.step_host_rows(bids)
        # ... and one aligned 16-byte record per unit (adc_step_out.unit_records)
        hu = d.step_host(bids, mode="records")
        obs, reward, term, trunc, _ = b.step({"keyword_bids": bids.cuda()})
        for hh in (h, hr, hu):
            for k in ("impressions", "buyside_clicks", "sellside_conversions"):
                assert torch.equal(hh[k].to(torch.int32), obs[k].cpu()), (k, step)
            for k in ("cost", "revenue"):
                assert torch.equal(hh[k], obs[k].cpu()), (k, step)
            assert torch.equal(hh["reward"], reward.cpu())
            assert torch.equal(hh["cumulative_profit"], obs["cumulative_profit"].cpu())
            assert torch.equal(hh["days_passed"], obs["days_passed"].cpu())
            assert torch.equal(hh["terminated"].bool(), term.cpu()) and torch.equal(hh["truncated"].bool(), trunc.cpu())
            assert int(hh["count_overflow"].sum()) == 0
        # the device-side observation of the pipelined env is up to date as well
        assert torch.equal(a._out["impressions"], obs["impressions"])
        assert torch.equal(d._out["impressions"], obs["impressions"]) and torch.equal(d._out["cost"], obs["cost"])
    assert int(h["impressions"].to(torch.int64).sum()) > 0
    if drift:
        pa, pb = a.keyword_params(), b.keyword_params()
        assert all(np.array_equal(pa[n], pb[n]) for n in ("vol_mean", "ctr", "cvr"))


def test_unit_records_other_kernel_families_and_count_overflow():
    """adc_step_out.unit_records beyond the free-running implicit kernels: explicit keywords get their
    records from the packing pass over the finished step; a count above 65535 (a day of 70 000
    auctions, walked by the exact serial kernel) is stored as 65535 with the record's flag set while the
    device arrays keep the exact value; a per-env budget tensor goes through the same call."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    E, K = 37, 10
    table = kwm.sample_random_keywords(K, np.random.default_rng(0))
    mk = lambda: VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1000.0, device="cuda", seed=2)
    a, b = mk(), mk()
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    for step in range(3):
        bids = torch.from_numpy(np.round(rng.uniform(0.01, 3.0, (E, K)), 2).astype(np.float32)).pin_memory()
        budget = torch.from_numpy(rng.uniform(5.0, 2000.0, E).astype(np.float32)).pin_memory()
        h = a.step_host(bids, budget, mode="records")
        obs, reward, term, trunc, _ = b.step({"keyword_bids": bids.cuda(), "budget": budget.cuda()})
        for k in ("impressions", "buyside_clicks", "sellside_conversions"):
            assert torch.equal(h[k].to(torch.int32), obs[k].cpu()), (k, step)
        for k in ("cost", "revenue"):
            assert torch.equal(h[k], obs[k].cpu()), (k, step)
        assert torch.equal(h["reward"], reward.cpu()) and int(h["count_overflow"].sum()) == 0
    assert int(h["buyside_clicks"].to(torch.int64).sum()) > 0
    # counts beyond uint16
    K2, E2 = 3, 4
    big = kwm.sample_implicit_keywords_from_quantiles(K2, np.random.default_rng(2), {"mean_volume": 64, "conversion_rate": 0.8})
    big.vol_mean[:] = [70000.0, 100.0, 66000.0]
    big.vol_std[:] = 0.0
    mk2 = lambda: VectorBiddingSimulation(E2, num_keywords=K2, keywords=big, budget=1e9, device="cuda", seed=5)
    a, b = mk2(), mk2()
    a.reset(); b.reset()
    bids = torch.full((E2, K2), 50.0, dtype=torch.float32).pin_memory()
    h = a.step_host(bids, mode="records")
    obs = b.step({"keyword_bids": bids.cuda()})[0]
    imp = obs["impressions"].cpu()
    assert int(imp[:, 0].min()) > 65535 and int(imp[:, 1].max()) <= 65535
    assert torch.equal(h["impressions"].to(torch.int32), imp.clamp(max=65535))
    assert torch.equal(h["count_overflow"].bool(), (imp > 65535) | (obs["buyside_clicks"].cpu() > 65535)
                       | (obs["sellside_conversions"].cpu() > 65535))
    assert torch.equal(a._out["impressions"], obs["impressions"])
    assert torch.equal(h["cost"], obs["cost"].cpu())


def test_device_side_explicit_keyword_sampling_matches_host_distributions():
    """SURVEY 8f-2, the default env's factory (gymnasium_kw_utils.py:113-156): per-env ExplicitKeyword
    sets drawn on the GPU follow the distributions of the host factory that reproduces the
    reference's draws bit for bit (two-sample KS per parameter), and an env steps on them."""
    from scipy.stats import ks_2samp
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    E, K = 128, 100
    g = torch.Generator(device="cuda").manual_seed(5)
    cols = kwm.sample_random_keywords_device(E, K, torch.device("cuda", 0), generator=g)
    host = kwm.sample_random_keywords(K * 32, np.random.default_rng(9))
    for n in kwm.PARAM_NAMES:
        a, b = cols[n].cpu().numpy().ravel(), getattr(host, n).ravel()
        assert ks_2samp(a, b).pvalue > 1e-4, n
    assert float(cols["vol_mean"].min()) >= 14 and float(cols["vol_mean"].max()) <= 29
    g2 = torch.Generator(device="cuda").manual_seed(5)
    again = kwm.sample_random_keywords_device(E, K, torch.device("cuda", 0), generator=g2)
    assert all(torch.equal(cols[n], again[n]) for n in kwm.PARAM_NAMES)
    env = VectorBiddingSimulation(E, num_keywords=K, device="cuda", seed=1, budget=1e6)
    env.install_device_keywords(cols, kind=kwm.EXPLICIT)
    obs = env.step({"keyword_bids": torch.full((E, K), 1.5, device="cuda")})[0]
    assert int(obs["impressions"].sum()) > 0 and float(obs["cost"].max()) <= 4.4 * 40


@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.float64])
def test_step_host_auto_routes_by_observation_dtype(obs_dtype):
    """step_host(mode="auto") with one rank per host: 16-byte unit records for float32 observations, the
    int32 / float64 zero-copy arrays for float64 ones (a record carries float32 money) -- either way the
    host observation equals the device step's."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(21)
    E, K = 50, 40
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 128, "conversion_rate": 0.8})
    mk = lambda: VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=500.0, device="cuda", seed=8,
                                         obs_dtype=obs_dtype)
    a, b = mk(), mk()
    a.reset(); b.reset()
    for step in range(3):
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2).astype(np.float32)).pin_memory()
        h = a.step_host(bids, mode="auto")
        obs, reward, term, trunc, _ = b.step({"keyword_bids": bids.cuda()})
        assert ("count_overflow" in h) == (obs_dtype == torch.float32)
        for k in ("impressions", "buyside_clicks", "sellside_conversions"):
            assert torch.equal(h[k].to(torch.int64), obs[k].cpu().to(torch.int64)), (k, step)
        for k in ("cost", "revenue"):
            assert torch.equal(h[k].contiguous(), obs[k].cpu()), (k, step)
        assert torch.equal(h["reward"].double().view(-1), reward.cpu().double().view(-1))
    assert int(obs["buyside_clicks"].sum()) > 0
