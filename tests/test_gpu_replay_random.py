"""GPU replay (tape mode) against the C oracle on random batched tapes, incl. ragged and empty
streams, binding budgets and the ndarray-aliasing double charge."""
import numpy as np
import pytest

from conftest import make_explicit_table, make_implicit_table, oracle_keywordset

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _random_tape(orc, rng, kind, K, vols, explicit_p=0.5):
    comp, ucl, ucv, rev, cost = [], [], [], [], []
    impr = np.zeros((K, 24), np.int32)
    for k in range(K):
        V = int(vols[k])
        comp.append(rng.integers(0, 160, V))
        if kind == orc.EXPLICIT:
            q = V // 24
            n = [V - 23 * q] + [q] * 23
            impr[k] = [rng.binomial(nn, explicit_p) for nn in n]
            slots = int(np.maximum(impr[k], 1).sum())
            cost.append(np.clip(rng.normal(2.4, 0.3, int(impr[k].sum())), 0, 4.4))
        else:
            slots = V
        ucl.append(rng.random(slots))
        ucv.append(rng.random(slots))
        rev.append(rng.integers(1, 300, slots))
    return orc.Tape.from_lists(vols, comp, ucl, ucv, rev,
                               impr=impr if kind == orc.EXPLICIT else None,
                               cost=cost if kind == orc.EXPLICIT else None)


@pytest.mark.parametrize("n_lanes", [0, 1])
@pytest.mark.parametrize("kind_name", ["implicit", "explicit"])
@pytest.mark.parametrize("alias", [False, True])
def test_replay_matches_oracle(orc, kind_name, alias, n_lanes):
    from adcraft_b200.tape import DeviceTape
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(17 + alias)
    kind = orc.IMPLICIT if kind_name == "implicit" else orc.EXPLICIT
    K, E = 11, 37
    table = make_implicit_table(rng, K, 60) if kind == orc.IMPLICIT else make_explicit_table(rng, K)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1000.0, device="cuda",
                                  budget_alias=alias, obs_dtype=torch.float64, autoreset=False, n_lanes=n_lanes)
    env.reset()
    budgets = rng.choice([0.0, 0.4, 2.5, 9.0, 30.0, 1e6], size=E)
    cum = np.zeros(E)
    for step in range(3):
        vols = rng.integers(0, 90, (E, K))
        vols[rng.random((E, K)) < 0.1] += 300  # more than one 128-auction trip
        vols[rng.random((E, K)) < 0.15] = 0  # empty units
        tapes = [_random_tape(orc, rng, kind, K, vols[e]) for e in range(E)]
        bids = np.round(rng.uniform(0.01, 1.6, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step_replay(
            {"keyword_bids": torch.from_numpy(bids).cuda(), "budget": torch.from_numpy(budgets).cuda()},
            DeviceTape.from_host(tapes, "cuda"))
        for e in range(E):
            kw = oracle_keywordset(orc, table)
            out = orc.step_replay(kw, np.rint(bids[e] * 100).astype(np.int32), float(budgets[e]), tapes[e],
                                  budget_alias=alias)
            for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                         ("sellside_conversions", "conversions")):
                assert np.array_equal(obs[a][e].cpu().numpy(), out[b]), (a, e, step)
            np.testing.assert_allclose(obs["cost"][e].cpu().numpy(), out["cost"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(obs["revenue"][e].cpu().numpy(), out["revenue"], rtol=1e-9, atol=1e-12)
            assert abs(float(reward[e]) - out["reward"]) < 1e-9 * (1 + out["cost"].sum() + out["revenue"].sum())
            rb = float(env._out["remaining_budget"][e])
            assert abs(rb - out["remaining_budget"]) < 1e-9 * (1 + abs(budgets[e]))
            cum[e] += out["reward"]
        np.testing.assert_allclose(obs["cumulative_profit"][:, 0].cpu().numpy(), cum, rtol=1e-9, atol=1e-9)
