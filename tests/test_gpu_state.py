"""State handling of VectorBiddingSimulation around the hot path: reseeding, mask swaps mid-episode,
stepping an env that lives on a device other than the current one, the float32 tie rule switch."""
import numpy as np
import pytest

from conftest import make_implicit_table

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

KEYS = ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue")


def _run(env, bids, n):
    out = []
    for _ in range(n):
        obs, reward, *_ = env.step({"keyword_bids": bids})
        out.append({k: obs[k].clone() for k in KEYS} | {"reward": reward.clone()})
    return out


def _same(a, b):
    return all(torch.equal(x[k], y[k]) for x, y in zip(a, b) for k in x)


def test_seeded_reset_replays_the_trajectory():
    """reset(seed=s) twice on one object, and on an object with another history: identical
    trajectories (the Philox step counter rewinds, the scratch parity does not)."""
    from adcraft_b200.vector_env import VectorBiddingSimulation
    cfg = {"mean_volume": 64, "conversion_rate": 0.5}
    E, K = 33, 17
    bids = torch.full((E, K), 0.7, device="cuda")
    a = VectorBiddingSimulation(E, keyword_config=cfg, num_keywords=K, device="cuda", budget=40.0)
    a.reset(seed=11)
    first = _run(a, bids, 3)          # odd number of calls: the parity differs at the second reset
    a.reset(seed=11)
    again = _run(a, bids, 3)
    assert _same(first, again)
    b = VectorBiddingSimulation(E, keyword_config=cfg, num_keywords=K, device="cuda", budget=40.0)
    b.reset(seed=5)
    _run(b, bids, 4)
    b.reset(seed=11)
    assert _same(first, _run(b, bids, 3))
    a.reset(seed=12)
    assert not _same(first, _run(a, bids, 3))


def test_unseeded_envs_do_not_share_a_trajectory():
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(0)
    table = make_implicit_table(rng, 9, 64)
    bids = torch.full((8, 9), 0.8, device="cuda")
    runs = []
    for _ in range(2):
        env = VectorBiddingSimulation(8, num_keywords=9, keywords=table, device="cuda", budget=1e5)
        env.reset()  # no seed anywhere: the Philox key comes from OS entropy through np_random
        runs.append(_run(env, bids, 2))
    assert not _same(runs[0], runs[1])


def test_mask_swap_mid_episode_keeps_the_drift(orc):
    """set_updater_mask with a mask of the same popcount in the middle of an episode: the new mask
    is what the next steps use, and the parameters drifted so far survive (env:105-112 only swaps
    the mask).  Checked against the oracle, whose mask is swapped at the same step."""
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(3)
    E, K = 21, 10
    table = make_implicit_table(rng, K, 48)
    m1 = [True, True, False, False, True, False, False, False, False, False]
    m2 = [True, False, True, False, False, True, False, False, False, False]
    # shared 1-D table at construction: the first set_updater_mask broadens it to per-env copies
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, device="cuda", seed=77, budget=1e5,
                                  obs_dtype=torch.float64)
    env.reset()
    env.set_updater_mask(m1)
    ob = orc.BatchOracle(table.kind, E, K, {n: getattr(table, n) for n in kwm.PARAM_NAMES}, seed=77,
                         budget=1e5, drift_mask=np.array(m1, np.uint8))
    bids = np.round(rng.uniform(0.3, 1.2, (E, K)), 2)
    tb = torch.from_numpy(bids).cuda()

    def both():
        obs = env.step({"keyword_bids": tb})[0]
        ref = ob.step(bids, n_threads=2)
        for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                     ("sellside_conversions", "conversions")):
            assert np.array_equal(obs[a].cpu().numpy(), ref[b]), a
        cur = env.keyword_params()
        for n in ("vol_mean", "ctr", "cvr"):
            assert np.array_equal(cur[n], ob.p[n]), n

    for _ in range(3):
        both()
    drifted = env.keyword_params()["ctr"].copy()
    assert not np.array_equal(drifted[:, 0], np.broadcast_to(table.ctr[0], (E,)))
    junk = [torch.empty(E * K, device="cuda") for _ in range(8)]  # churn the allocator around the swap
    env.set_updater_mask(m2)
    ob.mask[:] = np.array(m2, np.uint8)
    del junk
    assert np.array_equal(env.keyword_params()["ctr"], drifted)  # nothing was re-uploaded
    for _ in range(3):
        both()


def test_env_on_a_device_that_is_not_current():
    """The ABI launches on the calling thread's current device; the env makes its own device
    current around every call (and the library refuses a mismatch instead of faulting)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(1)
    table = make_implicit_table(rng, 12, 64)
    bids = np.round(rng.uniform(0.3, 1.2, (16, 12)), 2)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        env = VectorBiddingSimulation(16, num_keywords=12, keywords=table, device=dev, seed=9, budget=1e5)
        env.reset()
        obs = env.step({"keyword_bids": torch.from_numpy(bids).to(dev)})[0]
        assert obs["impressions"].device == torch.device(dev)
        outs.append({k: obs[k].cpu() for k in KEYS})
    assert all(torch.equal(outs[0][k], outs[1][k]) for k in KEYS)
    assert torch.cuda.current_device() == 0


def test_library_refuses_a_device_mismatch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ctypes as C
    from adcraft_b200 import _capi
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(1)
    table = make_implicit_table(rng, 4, 16)
    env = VectorBiddingSimulation(4, num_keywords=4, keywords=table, device="cuda:1", seed=9)
    env.reset()
    bids = torch.full((4, 4), 0.5, device="cuda:1")
    a = env._fill_args(bids, None, False)
    torch.cuda.set_device(0)
    rc = env._lib.adc_step_philox(C.byref(a), C.c_void_p(0))
    assert rc == -1 and b"current device" in env._lib.adc_last_error()


def test_f32_tie_rule_is_a_switch():
    """f32_ties only acts on float32 bids whose float32 value lies above the cent value (0.30f):
    there the impressions can only grow; float64 bids and bids like 0.29f are untouched."""
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(2)
    table = make_implicit_table(rng, 6, 128)
    table.p1[:] = 0.30
    table.p2[:] = 0.02  # competitor bids pile up around 30 cents: ties are frequent
    res = {}
    for ties in (False, True):
        for cents, dt in ((30, torch.float32), (29, torch.float32), (30, torch.float64)):
            env = VectorBiddingSimulation(64, num_keywords=6, keywords=table, device="cuda", seed=4, budget=1e6,
                                          f32_ties=ties)
            env.reset()
            obs = env.step({"keyword_bids": torch.full((64, 6), cents / 100.0, dtype=dt, device="cuda")})[0]
            res[(ties, cents, dt)] = obs["impressions"].clone()
    assert (res[(True, 30, torch.float32)] >= res[(False, 30, torch.float32)]).all()
    assert (res[(True, 30, torch.float32)] > res[(False, 30, torch.float32)]).any()
    assert torch.equal(res[(True, 29, torch.float32)], res[(False, 29, torch.float32)])
    assert torch.equal(res[(True, 30, torch.float64)], res[(False, 30, torch.float64)])
