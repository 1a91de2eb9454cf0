"""Shared auctions (BASELINE config 4): A bidders inside one auction world.
The oracle's shared step is pinned to the reference's nth_price_auction in
tests/test_oracle_vs_reference.py; here the CUDA kernels must equal the oracle bit for bit."""
import numpy as np
import pytest

from conftest import make_implicit_table, oracle_keywordset

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _bids(rng, W, A, K):
    b = np.round(rng.uniform(0.05, 1.5, (W, A, K)), 2)
    b[:, 1, 0] = b[:, 0, 0] = b[:, :, 0].max(axis=1) + 0.01      # tied leaders on keyword 0
    b[0, :, 1] = 0.4                                              # everybody ties
    return b


@pytest.mark.parametrize("n_lanes,force_serial", [(0, False), (-8, False), (8, False), (0, True)])
def test_shared_auction_matches_oracle(orc, n_lanes, force_serial):
    from adcraft_b200.multi_agent import SharedAuctionSimulation
    rng = np.random.default_rng(21)
    W, A, K = 5, 8, 7
    table = make_implicit_table(rng, K, 50)
    sim = SharedAuctionSimulation(A, W, num_keywords=K, keywords=table, budget=1e6, device="cuda", seed=0xABC,
                                  env_base=40, obs_dtype=torch.float64, autoreset=False, n_lanes=n_lanes)
    sim.reset()
    budgets = rng.choice([0.5, 3.0, 20.0, 1e6], size=(W, A))          # some bind: exact serial walk per bidder
    for step in range(3):
        bids = _bids(rng, W, A, K)
        obs, reward, term, trunc, _ = sim.step(torch.from_numpy(bids).cuda(), torch.from_numpy(budgets).cuda(),
                                               force_serial=force_serial)
        cents = np.rint(np.maximum(bids, 0.01) * 100).astype(np.int32)
        for w in range(W):
            for a in range(A):
                floor = np.delete(cents[w], a, axis=0).max(axis=0).astype(np.int32)
                out = orc.step_philox_shared(oracle_keywordset(orc, table), cents[w, a], floor, float(budgets[w, a]),
                                             seed=0xABC, world_id=40 + w, step=step)
                for name, key in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                                  ("sellside_conversions", "conversions")):
                    assert np.array_equal(obs[name][w, a].cpu().numpy(), out[key]), (name, w, a, step)
                np.testing.assert_allclose(obs["cost"][w, a].cpu().numpy(), out["cost"], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(obs["revenue"][w, a].cpu().numpy(), out["revenue"], rtol=1e-12, atol=1e-12)
                assert abs(float(reward[w, a]) - out["reward"]) < 1e-9
        imp = obs["impressions"].cpu().numpy()
        assert ((imp > 0).sum(axis=1) <= 1).all()                  # one winner per (world, keyword) at most
        assert not imp[:, :, 0].any() and not imp[0, :, 1].any()   # ties at the top win nothing


def test_single_bidder_world_is_the_plain_env():
    from adcraft_b200.multi_agent import SharedAuctionSimulation
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(4)
    W, K = 64, 20
    table = make_implicit_table(rng, K, 100)
    kw = dict(num_keywords=K, keywords=table, budget=1e6, device="cuda", seed=9, obs_dtype=torch.float64)
    shared, plain = SharedAuctionSimulation(1, W, **kw), VectorBiddingSimulation(W, **kw)
    shared.reset(); plain.reset()
    for _ in range(2):
        bids = torch.from_numpy(np.round(rng.uniform(0.05, 1.5, (W, 1, K)), 2)).cuda()
        so = shared.step(bids)[0]
        po = plain.step({"keyword_bids": bids[:, 0].contiguous()})[0]
        for k in ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue"):
            assert torch.equal(so[k][:, 0], po[k]), k
