"""BASELINE.json's larger configurations through size-independent properties: sampled envs of the
full-size launch are re-simulated by the C oracle (same seeds, global env ids) and must match
bit for bit; the exact-serial route must equal the fast route on the same launch."""
import numpy as np
import pytest

from conftest import oracle_keywordset

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _check_sampled_envs(orc, env, table, bids_row, sample, seed, step, budget, mask=None):
    obs = env._obs()
    for e in sample:
        kw = oracle_keywordset(orc, table, e if table.per_env else 0)
        out = orc.step_philox(kw, np.rint(bids_row * 100).astype(np.int32), budget, seed=seed,
                              env_id=int(e), step=step, lanes=False)
        for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                     ("sellside_conversions", "conversions")):
            assert np.array_equal(obs[a][e].cpu().numpy(), out[b]), (a, e)
        assert np.array_equal(env._out["cost_cents"][e].cpu().numpy(), out["cost_cents"])
        assert np.array_equal(env._out["revenue_cents"][e].cpu().numpy(), out["revenue_cents"])


def test_c3_shape_sparse_non_stationary(orc):
    """C3: 1000 keywords x 16384 envs, non-stationary sparse (vol 64, cvr 0.1, mask all True)."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(5)
    K, E, seed = 1000, 16384, 77
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 64, "conversion_rate": 0.1})
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e5, device="cuda", seed=seed,
                                  updater_mask=[True] * K)
    env.reset()
    bids_row = np.round(rng.uniform(0.3, 1.2, K), 2)
    bids = torch.from_numpy(np.tile(bids_row, (E, 1)).astype(np.float32)).cuda()
    sample = [0, 1, 4095, 8191, 16383]
    kws = {e: oracle_keywordset(orc, table) for e in sample}
    for step in range(2):
        env.step({"keyword_bids": bids})
        obs = env._obs()
        for e in sample:
            out = orc.step_philox(kws[e], np.rint(bids_row * 100).astype(np.int32), 1e5, seed=seed, env_id=e,
                                  step=step, lanes=False)
            assert np.array_equal(obs["impressions"][e].cpu().numpy(), out["impressions"]), (e, step)
            assert np.array_equal(obs["sellside_conversions"][e].cpu().numpy(), out["conversions"])
            assert np.array_equal(env._out["cost_cents"][e].cpu().numpy(), out["cost_cents"])
            orc.drift_apply(kws[e], np.ones(K, bool), orc.drift_philox(K, seed, e, step), table.vol_std)
        cur = env._kw_dev
        for e in sample:
            for n in ("vol_mean", "ctr", "cvr"):
                assert np.array_equal(cur[n][e].cpu().numpy(), getattr(kws[e], n)), (n, e, step)


def test_c5_shape_one_gpu_shard(orc):
    """C5 per-GPU shard: 10 000 keywords x 131 072 envs (1.3e9 units, ~50 GB of outputs).  One step;
    sampled envs against the oracle with their GLOBAL ids (this rank pretends to be rank 3 of 8)."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    free, _total = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~70 GB of free HBM")
    rng = np.random.default_rng(6)
    K, E, seed, rank = 10_000, 131_072, 0x5EED, 3
    table = kwm.sample_implicit_keywords_from_quantiles(K, rng, {"mean_volume": 128, "conversion_rate": 0.8})
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e7, device="cuda", seed=seed,
                                  env_base=rank * E)
    env.reset()
    bids_row = np.round(rng.uniform(0.3, 1.2, K), 2)
    bids = torch.from_numpy(bids_row.astype(np.float32)).cuda().expand(E, K).contiguous()
    obs, reward, term, trunc, _ = env.step({"keyword_bids": bids})
    torch.cuda.synchronize()
    for e in (0, 65_537, E - 1):
        out = orc.step_philox(oracle_keywordset(orc, table), np.rint(bids_row * 100).astype(np.int32), 1e7,
                              seed=seed, env_id=rank * E + e, step=0, lanes=False)
        assert np.array_equal(obs["impressions"][e].cpu().numpy(), out["impressions"]), e
        assert np.array_equal(obs["buyside_clicks"][e].cpu().numpy(), out["clicks"]), e
        assert np.array_equal(env._out["revenue_cents"][e].cpu().numpy(), out["revenue_cents"]), e
        assert abs(float(reward[e]) - out["reward"]) < 1e-6 * (1 + out["cost"].sum() + out["revenue"].sum())
    # conservation: reward == sum(revenue_cents - cost_cents) / 100 for every env
    diff = (env._out["revenue_cents"].sum(1) - env._out["cost_cents"].sum(1)).double() / 100.0 - reward
    assert float(diff.abs().max()) < 1e-6
    del env, obs, bids
    torch.cuda.empty_cache()


def test_multi_agent_independent_copies():
    """C4 with the reference's semantics (multi_agent/env.py:30-33): A independent bidders per world."""
    from adcraft_b200.multi_agent import MultiAgentBiddingSimulation, make_multi_flat
    ma = MultiAgentBiddingSimulation(8, 64, num_keywords=10, seed=3, device="cuda")
    obs, info = ma.reset(seed=3)
    assert set(obs) == set(range(8)) and obs[0].shape == (64, 52)
    act = {a: torch.cat([torch.full((64, 1), 1000.0), torch.full((64, 10), 0.5 + 0.05 * a)], 1).cuda() for a in range(8)}
    o, r, te, tr, _ = ma.step(act)
    assert o[7].shape == (64, 52) and r[0].shape == (64,) and te["__all__"].shape == (64,)
    # agents are independent copies: different keyword sets -> different outcomes
    assert not torch.equal(o[0], o[1])
    single = make_multi_flat(2, num_keywords=10, seed=3, device="cuda")
    o2, _ = single.reset(seed=3)
    assert o2[1].shape == (1, 52)
