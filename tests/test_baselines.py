"""Vectorised NaiveZeroMarginStrategy against the reference class on the same observations and the
same uniform draws (CPU; needs the reference tree) plus shape / invariants everywhere."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import ref_harness as rh  # noqa: E402


def _fake_obs(rng, K):
    imp = rng.integers(0, 60, K)
    clicks = np.minimum(imp, rng.integers(0, 25, K)) * (rng.random(K) < 0.7)
    conv = np.minimum(clicks, rng.integers(0, 12, K)) * (rng.random(K) < 0.6)
    rev = np.round(conv * rng.uniform(0.3, 1.5, K), 2)
    cost = np.round(clicks * rng.uniform(0.1, 0.8, K), 2)
    return dict(impressions=imp.astype(float), buyside_clicks=clicks.astype(float), cost=cost,
                sellside_conversions=conv.astype(float), revenue=rev)


def test_shapes_and_rampup():
    from adcraft_b200.baselines import VectorNaiveZeroMarginStrategy
    E, K = 3, 5
    pol = VectorNaiveZeroMarginStrategy(E, K, seed=1)
    zero = {k: torch.zeros(E, K) for k in ("buyside_clicks", "sellside_conversions", "revenue")}
    pol.update_all_caches({"keyword_bids": torch.full((E, K), 0.01)}, zero)
    a = pol.sample_action()
    assert a["keyword_bids"].shape == (E, K) and a["budget"].shape == (E,)
    assert torch.allclose(a["keyword_bids"], torch.full((E, K), 0.04, dtype=torch.float64))  # ramp-up step
    assert torch.all(a["budget"] == 100.0 * K)


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
def test_matches_reference_class():
    import importlib
    rh.load_reference()
    ie = importlib.import_module("adcraft.baselines.interpolated_expectations")
    from adcraft_b200.baselines import VectorNaiveZeroMarginStrategy
    rng = np.random.default_rng(0)
    K = 12
    ref = ie.NaiveZeroMarginStrategy(K, seed=123)
    mine = VectorNaiveZeroMarginStrategy(1, K)
    bids = np.full(K, 0.01)
    for step in range(25):
        obs = _fake_obs(rng, K)
        ref.update_all_caches({"keyword_bids": bids}, {k: v.copy() for k, v in obs.items()})
        mine.update_all_caches({"keyword_bids": torch.tensor(bids)[None]},
                               {k: torch.tensor(v)[None] for k, v in obs.items()})
        # the reference draws one uniform per keyword that has no revenue observation, in order
        state = ref.rng.bit_generator.state
        need = [i for i in range(K) if ref.caches[i]["num_rpc_obs"] < 1]
        with np.errstate(divide="ignore"):
            act_ref = ref.sample_action()
        ref.rng.bit_generator.state = state
        draws = ref.rng.random(len(need))
        u = np.ones(K)
        u[need] = draws
        act = mine.sample_action(uniforms=torch.tensor(u)[None])
        np.testing.assert_allclose(act["keyword_bids"][0].numpy(), act_ref["keyword_bids"], rtol=1e-6, atol=1e-9)
        assert abs(float(act["budget"][0]) - act_ref["budget"]) < 1e-9
        for i in range(K):
            c = ref.caches[i]
            assert abs(float(mine.num_rpc_obs[0, i]) - float(c["num_rpc_obs"])) < 1e-9, (step, i)
            assert abs(float(mine.num_sctr_obs[0, i]) - float(c["num_sctr_obs"])) < 1e-9, (step, i)
            assert abs(float(mine.ave_sctr[0, i]) - float(c["ave_sctr"])) < 1e-5
            assert abs(float(mine.ave_rpc[0, i]) - float(c["ave_rpc"])) < 1e-5
        bids = act_ref["keyword_bids"]


def test_interpolation_strategy_shapes_and_prior():
    from adcraft_b200.baselines import VectorNaiveInterpolationStrategy
    E, K = 2, 4
    pol = VectorNaiveInterpolationStrategy(E, K, seed=3)
    a = pol.sample_action()
    assert a["keyword_bids"].shape == (E, K) and a["budget"].shape == (E,)
    # nothing observed: mass only below max_observed(0.03) + bid_step -> bids in {0.01 .. 0.05}
    assert float(a["keyword_bids"].min()) >= 0.01 - 1e-12 and float(a["keyword_bids"].max()) <= 0.05 + 1e-12


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")
def test_interpolation_strategy_matches_reference_class():
    """Step-by-step against adcraft.baselines.interpolated_expectations.NaiveInterpolationStrategy:
    same observations, same previous bids, the uniforms its Generator.choice would consume."""
    import importlib
    rh.load_reference()
    ie = importlib.import_module("adcraft.baselines.interpolated_expectations")
    from adcraft_b200.baselines import VectorNaiveInterpolationStrategy
    rng = np.random.default_rng(1)
    K = 9
    ref = ie.NaiveInterpolationStrategy(K, seed=77)
    mine = VectorNaiveInterpolationStrategy(1, K)
    bids = np.full(K, 0.01)
    for step in range(60):
        obs = _fake_obs(rng, K)
        if step % 7 == 3:
            obs["buyside_clicks"][:] = 0; obs["sellside_conversions"][:] = 0; obs["revenue"][:] = 0; obs["cost"][:] = 0
        ref.update_all_caches({"keyword_bids": bids}, {k: v.copy() for k, v in obs.items()})
        mine.update_all_caches({"keyword_bids": torch.tensor(bids)[None]},
                               {k: torch.tensor(v)[None] for k, v in obs.items()})
        margins, costs = mine.expected_margins_and_costs()
        has = []
        for i in range(K):
            m_ref, c_ref = ref.get_expected_margin_from_cache(i)
            np.testing.assert_allclose(margins[0, i].numpy(), m_ref, rtol=1e-9, atol=1e-12, err_msg=f"margin {step} {i}")
            np.testing.assert_allclose(costs[0, i].numpy(), c_ref, rtol=1e-9, atol=1e-12, err_msg=f"cost {step} {i}")
            has.append(ref.get_profit_acquisition_function(np.array(m_ref), index=i) is not None)
        state = ref.rng.bit_generator.state
        act_ref = ref.sample_action()
        ref.rng.bit_generator.state = state
        u = np.zeros(K)
        u[np.array(has)] = ref.rng.random(int(np.sum(has)))
        act = mine.sample_action(uniforms=torch.tensor(u)[None])
        np.testing.assert_allclose(act["keyword_bids"][0].numpy(), act_ref["keyword_bids"], rtol=0, atol=1e-12)
        assert abs(float(act["budget"][0]) - act_ref["budget"]) < 1e-6 * max(1.0, abs(act_ref["budget"]))
        bids = np.asarray(act_ref["keyword_bids"], float)
    assert len({round(float(b), 2) for b in bids}) > 1   # the agent moved off the prior
