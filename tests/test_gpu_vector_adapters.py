"""RL-library vector front ends: protocol shape, autoreset bookkeeping and that the
numbers are the wrapped env's own (which the parity tests pin to the oracle)."""
import numpy as np
import pytest

from conftest import make_implicit_table

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(E, K, max_days, autoreset, seed=7):
    from adcraft_b200.vector_env import VectorBiddingSimulation
    table = make_implicit_table(np.random.default_rng(3), K, 40)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e6, max_days=max_days,
                                  device="cuda", seed=seed, autoreset=autoreset)
    return env


def test_gymnasium_vector_adapter_same_step_autoreset():
    from adcraft_b200.vector_adapters import GymnasiumVectorAdapter
    from adcraft_b200.wrappers import flat_observations
    E, K = 6, 5
    vec = GymnasiumVectorAdapter(_env(E, K, 3, True))
    ref = _env(E, K, 3, True)
    obs, info = vec.reset()
    ref.reset()
    assert obs.shape == (E, 5 * K + 2) and not obs.any() and "keyword_params" in info
    act = np.concatenate([np.full((E, 1), 1e6, np.float32), np.full((E, K), 0.8, np.float32)], axis=1)
    for day in range(1, 7):
        o, r, te, tr, infos = vec.step(act)
        ro, rr, rte, rtr, _ = ref.step({"keyword_bids": torch.full((E, K), 0.8, device="cuda"),
                                        "budget": torch.full((E,), 1e6, device="cuda")})
        want = flat_observations(ro).cpu().numpy()
        assert np.array_equal(r, rr.cpu().numpy()) and np.array_equal(te, rte.cpu().numpy())
        if day % 3 == 0:                      # max_days = 3: every env terminates together
            assert te.all() and infos["_final_observation"].all()
            assert all(np.array_equal(infos["final_observation"][i], want[i]) for i in range(E))
            assert not o.any()                # the returned observation is the reset one
        else:
            assert not te.any() and "final_observation" not in infos
            assert np.array_equal(o, want)
            assert o[0, 2 * K + 1] == day % 3  # days_passed slot of the sorted layout


def test_sb3_vecenv_adapter_protocol():
    from adcraft_b200.vector_adapters import SB3VecEnvAdapter
    E, K = 4, 3
    vec = SB3VecEnvAdapter(_env(E, K, 2, True))
    assert vec.observation_space.shape == (5 * K + 2,) and vec.action_space.shape == (K + 1,)
    assert vec.seed(5) == [5, 6, 7, 8]
    obs = vec.reset()
    assert obs.shape == (E, 5 * K + 2)
    act = np.concatenate([np.full((E, 1), 1e6, np.float32), np.full((E, K), 0.6, np.float32)], axis=1)
    o1, r1, d1, i1 = vec.step(act)
    assert not d1.any() and i1 == [{} for _ in range(E)]
    vec.step_async(act)
    o2, r2, d2, i2 = vec.step_wait()
    assert d2.all() and all("terminal_observation" in i and i["TimeLimit.truncated"] is False for i in i2)
    assert i2[0]["terminal_observation"][2 * K + 1] == 2 and not o2.any()
    assert vec.get_attr("num_keywords") == [K] * E and vec.env_is_wrapped(object) == [False] * E
    assert vec.get_attr("max_days", indices=[1, 2]) == [2, 2]


def test_rllib_vector_adapter_reset_at():
    from adcraft_b200.vector_adapters import RLlibVectorAdapter
    E, K = 3, 4
    vec = RLlibVectorAdapter(_env(E, K, 2, False))
    obs, infos = vec.vector_reset()
    assert len(obs) == E and obs[0].shape == (5 * K + 2,)
    act = [np.concatenate([[1e6], np.full(K, 0.9)]).astype(np.float32) for _ in range(E)]
    vec.vector_step(act)
    o, r, te, tr, _ = vec.vector_step(act)
    assert all(te) and o[1][2 * K + 1] == 2
    o1, _ = vec.reset_at(1)                    # only sub-env 1 starts over
    assert not o1.any()
    o, r, te, tr, _ = vec.vector_step(act)
    assert o[1][2 * K + 1] == 1 and o[0][2 * K + 1] == 3 and not te[1] and te[0]


@pytest.mark.parametrize("kind,budget", [("implicit", 1e5), ("implicit", 20.0), ("explicit", 1000.0), ("explicit", 5.0)])
def test_kernel_written_flat_rows_equal_the_gathered_layout(kind, budget):
    """adc_step_out.flat_obs: the [E, 5K+2] rows the kernels write (fast path, exact serial walk,
    explicit keywords) are the FlatArrayWrapper layout of the same step's observation dict
    (wrappers/flat_array.py:44-87; keys sorted, gymnasium_kw_utils.py:383-390)."""
    from conftest import make_explicit_table, make_implicit_table
    from adcraft_b200.vector_env import VectorBiddingSimulation
    from adcraft_b200.wrappers import flat_observations, observation_slices
    rng = np.random.default_rng(12)
    E, K = 37, 11
    table = make_implicit_table(rng, K, 64) if kind == "implicit" else make_explicit_table(rng, K)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, device="cuda", seed=3, budget=budget, max_days=2,
                                  flat_obs=True, obs_dtype=torch.float64)
    obs, _ = env.reset()
    assert float(env.flat_observation().abs().sum()) == 0.0
    for step in range(4):
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 2.5, (E, K)), 2)).cuda()
        obs, reward, term, trunc, _ = env.step({"keyword_bids": bids})
        flat = flat_observations(obs, env)
        assert flat.data_ptr() == env.flat_observation().data_ptr() and flat.shape == (E, 5 * K + 2)
        assert torch.equal(flat, flat_observations(obs)), step  # the gathered form of the same dict
    sl = observation_slices(K)
    assert torch.equal(flat[:, sl["days_passed"]], obs["days_passed"].to(flat.dtype))
    assert float(flat[:, sl["impressions"]].sum()) > 0
