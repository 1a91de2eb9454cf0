"""Pins the C oracle to the UNMODIFIED reference Python, live (build container only; skipped on
the GPU box, where /root/reference does not exist -- the committed goldens cover it there)."""
import tempfile

import numpy as np
import pytest

from oracle import ref_harness as rh

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")]


def _check(a, b, name):
    for f in ("impressions", "clicks", "conversions", "lane_I", "lane_B", "lane_S"):
        assert np.array_equal(np.asarray(a[f], np.int64), np.asarray(b[f], np.int64)), (name, f)
    for f in ("cost", "revenue", "profit"):
        assert np.array_equal(np.asarray(a[f]), np.asarray(b[f])), (name, f)
    assert a["reward"] == b["reward"] and a["lanes_run"] == b["lanes_run"], name


@pytest.mark.parametrize("vol,cvr,K,budget,seed", [
    (16, 0.5, 2, 1000.0, 0), (128, 0.8, 7, np.array([100000.0]), 1), (64, 0.1, 5, np.array([30.0]), 2),
    (128, 0.8, 6, 55.5, 3), (16, 0.1, 9, np.array([2.0]), 4), (64, 0.8, 4, 0.0, 5)])
def test_recorded_implicit_steps(orc, vol, cvr, K, budget, seed):
    from oracle import ref_driver as rd
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    env = ref["env"].bidding_sim_creator(dict(
        keyword_config=rh.experiment_keyword_config(vol, cvr, tmp), num_keywords=K, max_days=10,
        updater_mask=[True] * K))
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    for s in range(3):
        bids = np.round(rng.uniform(0.2, 1.5, size=K), 2)
        r = rd.record_step(env, {"keyword_bids": bids, "budget": budget})
        o = orc.step_replay(r["kw_before"], r["bid_cents"], r["budget"], r["tape"], budget_alias=r["budget_alias"])
        _check(r, o, (vol, K, s))


@pytest.mark.parametrize("K,budget,seed", [(1, 1000.0, 0), (10, np.array([1000.0]), 1), (10, np.array([20.0]), 2),
                                           (4, 3.0, 3)])
def test_recorded_explicit_steps(orc, K, budget, seed):
    from oracle import ref_driver as rd
    ref = rh.load_reference()
    env = ref["env"].BiddingSimulation(num_keywords=K)
    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    for s in range(3):
        bids = np.round(rng.uniform(0.01, 3.0, size=K), 2)
        r = rd.record_step(env, {"keyword_bids": bids, "budget": budget})
        o = orc.step_replay(r["kw_before"], r["bid_cents"], r["budget"], r["tape"], budget_alias=r["budget_alias"])
        _check(r, o, ("explicit", K, s))


@pytest.mark.parametrize("kind_name", ["implicit", "explicit"])
@pytest.mark.parametrize("K,vol,budget,alias", [(3, 16, 1000.0, False), (8, 128, 1e5, False), (6, 64, 20.0, False),
                                                (6, 64, 20.0, True), (5, 128, 3.5, True)])
def test_philox_tapes_through_the_reference(orc, kind_name, K, vol, budget, alias):
    """Free-running oracle -> recorded tape -> unmodified reference: same integers, same floats,
    same drifted parameters, same cumulative profit."""
    from conftest import make_explicit_table, make_implicit_table, oracle_keywordset
    from oracle import ref_driver as rd
    rng = np.random.default_rng(K * 100 + vol)
    table = make_implicit_table(rng, K, vol) if kind_name == "implicit" else make_explicit_table(rng, K)
    kw = oracle_keywordset(orc, table)
    mask = np.ones(K, bool)
    env = rd.build_replay_env(kw, budget=budget, drift_mask=mask)
    kwc, cum = kw.copy(), 0.0
    for step in range(3):
        bids = np.round(rng.uniform(0.2, 1.5 if kind_name == "implicit" else 3.0, K), 2)
        bc = np.rint(bids * 100).astype(np.int32)
        o = orc.step_philox(kwc, bc, budget, seed=1234, env_id=5, step=step, record_cap=4096, budget_alias=alias)
        tape = o["tape"]
        tape.drift = orc.drift_philox(K, 1234, 5, step)
        r = rd.replay_step(env, bids, np.array([budget]) if alias else None, tape)
        _check(r, o, (kind_name, K, step))
        orc.drift_apply(kwc, mask, tape.drift, kw.vol_std)
        for n in ("vol_mean", "ctr", "cvr"):
            assert np.array_equal(getattr(r["kw_after"], n), getattr(kwc, n)), n
        cum += o["reward"]
        assert r["cumulative_profit"] == cum


def test_keyword_factories_match_reference_rng_order(orc):
    from adcraft_b200 import keywords as kwm
    from oracle import ref_driver as rd
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    for vol, cvr, K, seed in [(16, 0.5, 2, 0), (100, 0.3, 30, 10), (128, 0.8, 100, 5)]:
        env = ref["env"].bidding_sim_creator(dict(keyword_config=rh.experiment_keyword_config(vol, cvr, tmp),
                                                  num_keywords=K))
        env.reset(seed=seed)
        a = rd.keywordset_from_env(env)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        t = kwm.sample_implicit_keywords_from_quantiles(K, g, {"mean_volume": vol, "conversion_rate": cvr})
        for n in kwm.PARAM_NAMES:
            np.testing.assert_allclose(getattr(a, n), getattr(t, n), rtol=1e-15 if n == "p2" else 0)
        assert env.np_random.random() == g.random()  # generator left in the same state
    for K, seed in [(10, 1), (1, 0)]:
        env = ref["env"].BiddingSimulation(num_keywords=K)
        env.reset(seed=seed)
        a = rd.keywordset_from_env(env)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        t = kwm.sample_random_keywords(K, g)
        for n in kwm.PARAM_NAMES:
            assert np.array_equal(getattr(a, n), getattr(t, n)), n
        assert env.np_random.random() == g.random()


def test_shared_auction_is_nth_price_auction_on_rivals_plus_competitor(orc):
    """with A bidders in one auction, bidder a's ``other_bids`` are the A-1 rival
    bids plus the keyword's sampled competitor, cleared by the reference's own
    ``nth_price_auction(n=2, num_winners=1)`` (synthetic_kw_helpers.py:116-180).  The oracle's
    shared step (clearing price = max(competitor, highest rival)) must give the same impressions
    and costs for every bidder, ties included."""
    ref = rh.load_reference()
    npa = ref["helpers"].nth_price_auction
    rng = np.random.default_rng(8)
    K, A = 6, 8
    loc = rng.uniform(0.3, 1.0, K)
    kw = orc.KeywordSet(orc.IMPLICIT, np.full(K, 60.0), np.full(K, 9.0), loc, 0.2 * loc, np.ones(K),  # ctr 1:
                        rng.uniform(0.1, 0.9, K), rng.uniform(0.3, 1.5, K), rng.uniform(0.01, 0.3, K))  # all clicks
    for step in range(3):
        bids = rng.integers(20, 140, (A, K)).astype(np.int32)
        bids[1, 0] = bids[0, 0] = bids[:, 0].max() + 1          # a tie at the top: nobody wins keyword 0
        bids[3, 1] = 1                                           # a bidder far below the field
        winners = np.zeros(K, int)
        for a in range(A):
            # the world's draws as bidder a sees them without rivals: the free-running mode decides
            # every auction by one shared uniform R_j (win <=> R_j < T1(bid)), so the recorded
            # competitor stream of a solo run holds the price of every auction this bid would win
            # and a losing value everywhere else; all bidders of a world share the R_j
            solo = orc.step_philox(kw, bids[a], 1e9, seed=5, env_id=2, step=step, record_cap=512)
            comp = [np.asarray(solo["tape"].comp_cents[solo["tape"].comp_off[k]:solo["tape"].comp_off[k + 1]])
                    for k in range(K)]
            rivals = np.delete(bids, a, axis=0)
            floor = rivals.max(axis=0).astype(np.int32)
            out = orc.step_philox_shared(kw, bids[a], floor, 1e9, seed=5, world_id=2, step=step)
            for k in range(K):
                V = len(comp[k])
                other = np.hstack([np.tile(rivals[:, k] / 100.0, (V, 1)), comp[k][:, None] / 100.0])
                I, _, costs = npa(bids[a, k] / 100.0, other, n=2, num_winners=1)
                assert out["impressions"][k] == I, (step, a, k)
                assert abs(out["cost"][k] - float(np.sum(costs))) < 1e-9, (step, a, k)
                winners[k] += I > 0
        assert winners[0] == 0 and (winners <= 1).all()


def test_quantile_csv_wire_format_against_reference(orc):
    """The quantile CSVs: a file written by the reference (pandas to_csv) reads back
    to the same table here, and a multi-bucket file written here (a bucket without data for one
    parameter, a NaN-volume bucket) gives the reference's sampler -- through its own default loader,
    gymnasium_kw_utils.py:238-257 -- exactly the keywords it gives ours, generator state included."""
    import os
    from adcraft_b200 import keywords as kwm
    from oracle import ref_driver as rd
    ref = rh.load_reference()
    pytest.importorskip("pandas")
    tmp = tempfile.mkdtemp()
    # (1) reference-written experiment file -> our reader
    cfg = rh.experiment_keyword_config(64, 0.1, tmp)
    cfg["make_quant_func"](cfg)
    theirs = cfg["load_quant_func"](cfg)
    ours = kwm.load_experiment_quantiles(cfg)
    assert set(ours) == {c for c in theirs.columns if not c.startswith("Unnamed")}
    for c in ours:
        np.testing.assert_array_equal(ours[c], np.asarray(theirs[c], dtype=np.float64))
    # (2) our multi-bucket file -> the reference's default loader and sampler
    rng = np.random.default_rng(1)
    rows, cols = 5, {}
    for p in kwm.QUANTILE_PARAMS:
        lo = np.round(rng.uniform(0.05, 0.4, rows), 3)
        # (the bucket without sctr data is the LAST one: the reference indexes the filtered pandas
        # Series by label, quantiles_to_keywords.py:24-26, so a gap in the middle raises KeyError there)
        cols[f"count_{p}"] = np.array([3.0, 1.0, 5.0, 2.0, 0.0]) if p == "sctr" else np.full(rows, 4.0)
        cols[f"min_{p}"], cols[f"median_{p}"], cols[f"max_{p}"] = lo, lo + 0.125, lo + 0.5
    cols["min_vol"] = np.array([8.0, 20, np.nan, 90, 300])
    cols["median_vol"] = np.array([12.0, 40, np.nan, 120, 400])
    cols["max_vol"] = np.array([16.0, 64, np.nan, 256, 500])
    os.makedirs(os.path.join(tmp, "prod"))
    kwm.write_quantile_csv(cols, os.path.join(tmp, "prod", "auction_data.csv"))
    kcfg = {"outer_directory": tmp + "/", "quantiles_folder": "prod/", "no_vol_prob": 0.2}
    for K, seed in [(3, 0), (40, 7)]:
        env = ref["env"].bidding_sim_creator(dict(keyword_config=dict(kcfg), num_keywords=K))
        env.reset(seed=seed)
        a = rd.keywordset_from_env(env)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        t = kwm.sample_implicit_keywords_from_quantiles(K, g, dict(kcfg))
        for n in kwm.PARAM_NAMES:
            np.testing.assert_allclose(getattr(a, n), getattr(t, n), rtol=1e-15 if n == "p2" else 0, err_msg=n)
        assert env.np_random.random() == g.random()


def _multi_set(orc, rng, K):
    return orc.KeywordSet(orc.IMPLICIT_MULTI, np.full(K, 40.0), np.full(K, 6.0), rng.uniform(-0.1, 0.2, K),
                          rng.uniform(0.05, 0.2, K), rng.uniform(0.2, 0.9, K), rng.uniform(0.2, 0.9, K),
                          rng.uniform(0.3, 1.5, K), rng.uniform(0.05, 0.3, K),
                          max_bidders=rng.choice([0.0, 1.0, 2.0, 3.0, 5.0, 30.0], K), participation=rng.uniform(0.2, 0.9, K))


@pytest.mark.parametrize("budget,seed", [(1000.0, 0), (np.array([4.0]), 1), (2.0, 2)])
def test_recorded_multi_bidder_steps(orc, budget, seed):
    """The class-default ImplicitKeyword (classes:578-688: Binomial bidders per lane, signed Laplace
    bids, nth_price_auction with zero padding) on its own numpy Generator -> tape -> oracle replay."""
    from oracle import ref_driver as rd
    rng = np.random.default_rng(seed)
    kw = _multi_set(orc, rng, 6)
    env = rd.build_multi_env(kw, seed=seed)
    for s in range(3):
        bids = np.round(rng.uniform(0.01, 0.5, 6), 2)
        r = rd.record_step(env, {"keyword_bids": bids, "budget": budget})
        o = orc.step_replay(r["kw_before"], r["bid_cents"], r["budget"], r["tape"], budget_alias=r["budget_alias"])
        _check(r, o, ("multi", s))
    assert (np.asarray(r["tape"].impr) < 3).any() and (np.asarray(r["tape"].impr) >= 3).any()


@pytest.mark.parametrize("budget,alias", [(1000.0, False), (3.0, False), (5.0, True)])
def test_philox_multi_bidder_tapes_through_the_reference(orc, budget, alias):
    from oracle import ref_driver as rd
    rng = np.random.default_rng(11)
    kw = _multi_set(orc, rng, 5)
    env = rd.build_replay_env(kw, budget=budget)
    for step in range(3):
        bids = np.round(rng.uniform(0.02, 0.5, 5), 2)
        bc = np.rint(bids * 100).astype(np.int32)
        o = orc.step_philox(kw, bc, budget, seed=77, env_id=3, step=step, record_cap=4096, budget_alias=alias)
        r = rd.replay_step(env, bids, np.array([budget]) if alias else None, o["tape"])
        _check(r, o, ("multi-philox", step))
