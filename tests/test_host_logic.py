"""Host-side pieces that need no GPU: flat views, metrics, sharding, spaces."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_flat_observation_order_and_slices():
    from adcraft_b200 import wrappers as w
    K = 3
    obs = dict(impressions=np.array([1, 2, 3]), buyside_clicks=np.array([4, 5, 6]), cost=np.array([.1, .2, .3]),
               sellside_conversions=np.array([7, 8, 9]), revenue=np.array([1.5, 2.5, 3.5]),
               cumulative_profit=np.array([-4.0]), days_passed=np.array([2]))
    flat = w.flatten_dict_array(obs)
    # gymnasium_kw_utils.py:383-390: keys sorted; RL/train_agent.ipynb slices rely on this layout
    assert np.allclose(flat, [4, 5, 6, .1, .2, .3, -4.0, 2, 1, 2, 3, 1.5, 2.5, 3.5, 7, 8, 9])
    sl = w.observation_slices(K)
    for k in obs:
        assert np.allclose(flat[sl[k]], np.asarray(obs[k], float).ravel())
    tobs = {k: torch.as_tensor(np.asarray(v))[None].repeat(2, *([1] * np.asarray(v).ndim)) for k, v in obs.items()}
    tf = w.flat_observations({k: v.double() if v.dtype.is_floating_point else v for k, v in tobs.items()})
    assert tf.shape == (2, 5 * K + 2) and np.allclose(tf[1].numpy(), flat)
    act = w.unflatten_actions(torch.arange(8.0).reshape(2, 4))
    assert act["budget"].tolist() == [0.0, 4.0] and act["keyword_bids"].shape == (2, 3)


def test_metrics_definitions():
    from adcraft_b200 import metrics as m
    rng = np.random.default_rng(0)
    kw, ideal = rng.normal(1, 1, (60, 7)), np.abs(rng.normal(2, 1, (60, 7)))
    ideal[:, 3] = 0.0  # non-positive ideal -> denominator 1 (experiment_metrics.py:70-72)
    den = ideal.copy(); den[den <= 0] = 1.0
    assert m.compute_AKNCP(kw, ideal) == float(np.median(kw.mean(0) / den.mean(0)))
    assert m.compute_NCP(kw, ideal) == float(kw.sum() / ideal.sum())
    assert m.compute_NCP(kw, np.zeros_like(ideal)) == float(kw.sum())
    acc = m.MetricAccumulator(1, 7, "cpu")
    for t in range(60):
        obs = {"revenue": torch.tensor(kw[t][None]), "cost": torch.zeros(1, 7, dtype=torch.float64)}
        acc.update(obs, torch.tensor([kw[t].sum()]), ideal=torch.tensor(ideal[t][None]))
    pe = acc.per_env()
    assert abs(float(pe["akncp"][0]) - m.compute_AKNCP(kw, ideal)) < 1e-12
    assert abs(float(pe["ncp"][0]) - m.compute_NCP(kw, ideal)) < 1e-12
    s = m.summarize(acc.summary_vector())
    assert s["n_envs"] == 1.0 and abs(s["reward_sum"] - kw.sum()) < 1e-9


def test_env_range_partitions_everything():
    from adcraft_b200.sharding import env_range
    for total, world in [(4096, 8), (10, 3), (7, 8), (1 << 20, 8)]:
        spans = [env_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_spaces_match_reference_definitions():
    from adcraft_b200.spaces import get_action_space, get_observation_space
    a, o = get_action_space(4), get_observation_space(4, 1000.0)
    assert a["keyword_bids"].shape == (4,) and a["keyword_bids"].dtype == np.float32
    assert a["budget"].shape == (1,)
    assert set(o.spaces.keys()) == {"impressions", "buyside_clicks", "cost", "sellside_conversions", "revenue",
                                    "cumulative_profit", "days_passed"}
    zero = dict(impressions=np.zeros(4, np.int64), buyside_clicks=np.zeros(4, np.int64), cost=np.zeros(4, np.float32),
                sellside_conversions=np.zeros(4, np.int64), revenue=np.zeros(4, np.float32),
                cumulative_profit=np.zeros(1, np.float32), days_passed=np.zeros(1, np.float32))
    assert o.contains(zero)  # tests/test_env.py:52-57
    assert a.contains(a.sample())


def test_rust_float_formatting():
    """info["bidding_outcomes"] is built by Rust's format! (src/lib.rs:269): `{}` on f64 is always
    positional and drops a trailing .0; `{:?}` (inside Vec<f64>) keeps one fractional digit and
    switches to `1.5e-7`-style exponents below 1e-4 and from 1e16 on."""
    from adcraft_b200.gymnasium_kw_env import _rust_debug, _rust_display, repr_outcomes
    assert [_rust_display(v) for v in (0.75, 1.0, 0.0, 100.0, 1e-7, -2.5e-16, 1e21, 0.30000001192092896)] == [
        "0.75", "1", "0", "100", "0.0000001", "-0.00000000000000025", "1000000000000000000000",
        "0.30000001192092896"]
    assert [_rust_debug(v) for v in (0.57, 1.0, 0.0, 1e-4, 9.9e-5, 1.5e-7, 1e16, 123456.789)] == [
        "0.57", "1.0", "0.0", "0.0001", "9.9e-5", "1.5e-7", "1e16", "123456.789"]
    s = repr_outcomes([dict(bid=0.5, impressions=3, impression_share=0.6, buyside_clicks=2, costs=[0.41, 0.0],
                            sellside_conversions=1, revenues=[1.0], revenues_per_cost=[0.0, 1.0], profit=0.59)])
    assert s == ("[{'bid': 0.5, 'impressions': 3, 'impression_share': 0.6, 'buyside_clicks': 2, "
                 "'costs': [0.41, 0.0], 'sellside_conversions': 1, 'revenues': [1.0], "
                 "'revenues_per_cost': [0.0, 1.0], 'profit': 0.59}]")
