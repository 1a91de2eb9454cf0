"""GPU parity: the CUDA path (through the C ABI) against the C oracle on the same inputs.

Integer outcomes (impressions, clicks, conversions, flags, exact cents) must be bit-exact;
float outputs within 1e-6 relative in f64 (1e-4 in f32) as BASELINE.json's north_star states
(in practice the f64 results are exact: money is integer cents on both sides).
"""
import numpy as np
import pytest

from conftest import make_explicit_table, make_implicit_table, oracle_keywordset

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-6, 1e-4


def _env(table, E, **kw):
    from adcraft_b200.vector_env import VectorBiddingSimulation
    env = VectorBiddingSimulation(E, num_keywords=table.K, keywords=table, device="cuda", **kw)
    env.reset()
    return env


def _oracle_batch(orc, table, E, *, seed, budget, mask=None, alias=False, max_days=60,
                  loss_threshold=10000.0, step0=0, env_base=0):
    from adcraft_b200 import keywords as kwm
    params = {n: getattr(table, n) for n in kwm.PARAM_NAMES}
    return orc.BatchOracle(table.kind, E, table.K, params, seed=seed, budget=budget, max_days=max_days,
                           loss_threshold=loss_threshold, drift_mask=mask, env_base=env_base,
                           step0=step0, budget_alias=alias, impression_thresh=table.impression_thresh)


def _compare(obs, reward, term, trunc, ref, env, rtol):
    for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                 ("sellside_conversions", "conversions")):
        got = obs[a].cpu().numpy()
        assert np.array_equal(got, ref[b]), f"{a} differs at {np.argwhere(got != ref[b])[:5]}"
    assert np.array_equal(term.cpu().numpy().astype(np.uint8), ref["terminated"])
    assert np.array_equal(trunc.cpu().numpy().astype(np.uint8), ref["truncated"])
    for a in ("cost", "revenue"):
        np.testing.assert_allclose(obs[a].cpu().numpy().astype(np.float64), ref[a], rtol=rtol, atol=1e-9)
    scale = np.abs(ref["cost"]).sum(1) + np.abs(ref["revenue"]).sum(1) + 1.0
    assert np.all(np.abs(reward.cpu().numpy() - ref["reward"]) <= rtol * scale)


@pytest.mark.parametrize("n_lanes", [0, -8, -16, 1, 4, 8, 32])
@pytest.mark.parametrize("vol,K,E,budget", [(128, 100, 64, 1e5), (16, 37, 50, 1e5), (300, 5, 33, 1e5),
                                            (64, 20, 40, 25.0), (0, 3, 4, 10.0)])
def test_philox_implicit_matches_oracle(orc, n_lanes, vol, K, E, budget):
    rng = np.random.default_rng(vol * 1000 + K)
    table = make_implicit_table(rng, K, max(vol, 1))
    if vol == 0:
        table.vol_mean[:] = 0.0
        table.vol_std[:] = 0.3
    env = _env(table, E, seed=4242, budget=budget, n_lanes=n_lanes, obs_dtype=torch.float64, max_days=3)
    ob = _oracle_batch(orc, table, E, seed=4242, budget=budget, max_days=3)
    for s in range(5):
        bids = np.round(rng.uniform(0.05, 1.6, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
        assert np.array_equal(env._out["cost_cents"].cpu().numpy(), np.rint(ref["cost"] * 100).astype(np.int64))


@pytest.mark.parametrize("alias", [False, True])
def test_budget_binding_serial_path(orc, alias):
    rng = np.random.default_rng(7)
    K, E = 12, 48
    table = make_implicit_table(rng, K, 96)
    budgets = rng.choice([0.5, 3.0, 12.0, 40.0, 1e5], size=E)
    env = _env(table, E, seed=99, budget=1000.0, budget_alias=alias, obs_dtype=torch.float64)
    ob = _oracle_batch(orc, table, E, seed=99, budget=1000.0, alias=alias)
    for s in range(4):
        bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
        ob.budget[:] = budgets
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda(),
                                                "budget": torch.from_numpy(budgets).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)


@pytest.mark.parametrize("K,E,vol", [(150, 12, 40), (2300, 5, 12)])
def test_budget_binding_serial_path_large_keyword_sets(orc, K, E, vol):
    """The exact serial kernel keeps the per-unit constants of the first 104 keywords in shared
    memory and recomputes the rest every sub-step: both tiers, with budgets that bind at different
    sub-steps (and not at all for some envs), must equal the oracle."""
    rng = np.random.default_rng(17)
    table = make_implicit_table(rng, K, vol)
    day_spend = K * vol * 0.15  # rough scale of a day's spend in dollars
    budgets = np.resize([0.02, 0.2, 0.6, 5.0], E) * day_spend
    env = _env(table, E, seed=31, budget=1000.0, obs_dtype=torch.float64)
    ob = _oracle_batch(orc, table, E, seed=31, budget=1000.0)
    for s in range(2):
        bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
        ob.budget[:] = budgets
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda(),
                                                "budget": torch.from_numpy(budgets).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
    assert float(obs["cost"].sum()) > 0


@pytest.mark.parametrize("alias", [False, True])
def test_serial_hint_changes_nothing(orc, alias):
    """adc_scratch.serial_hint: envs whose budget bound skip the budget-free kernel in their next step.
    Same results with and without it, budgets that start / stop binding between steps included, and the
    marks are exactly the envs whose budget could not cover the day."""
    rng = np.random.default_rng(23)
    K, E = 40, 96
    table = make_implicit_table(rng, K, 100)
    a = _env(table, E, seed=77, budget=1000.0, budget_alias=alias, obs_dtype=torch.float64)
    b = _env(table, E, seed=77, budget=1000.0, budget_alias=alias, obs_dtype=torch.float64, serial_hint=False)
    ob = _oracle_batch(orc, table, E, seed=77, budget=1000.0, alias=alias)
    marked = 0
    for s in range(6):
        bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
        budgets = rng.choice([2.0, 30.0, 300.0, 1e6], size=E)
        act = {"keyword_bids": torch.from_numpy(bids).cuda(), "budget": torch.from_numpy(budgets).cuda()}
        oa, ob_ = a.step(act), b.step(act)
        for k in oa[0]:
            assert torch.equal(oa[0][k], ob_[0][k]), (s, k)
        assert torch.equal(oa[1], ob_[1]) and torch.equal(oa[2], ob_[2])
        ob.budget[:] = budgets
        _compare(oa[0], oa[1], oa[2], oa[3], ob.step(bids, n_threads=4), a, RTOL64)
        hint = a._scratch["serial_hint"].cpu().numpy()
        assert not hint[budgets >= 1e6].any()     # a budget that covers everything leaves no mark
        assert hint[budgets <= 2.0].all()         # a $2 budget cannot cover a day of 40 keywords
        marked += int(hint.sum())
        assert not b._scratch["serial_hint"].any()
    assert marked > 0


def test_force_serial_equals_fast_path(orc):
    rng = np.random.default_rng(11)
    K, E = 30, 40
    table = make_implicit_table(rng, K, 128)
    bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2)).cuda()
    a = _env(table, E, seed=5, budget=1e6, obs_dtype=torch.float64)
    b = _env(table, E, seed=5, budget=1e6, obs_dtype=torch.float64)
    for s in range(3):
        oa = a.step({"keyword_bids": bids})
        ob_ = b.step({"keyword_bids": bids}, force_serial=True)
        for k in oa[0]:
            assert torch.equal(oa[0][k], ob_[0][k]), k
        assert torch.equal(oa[1], ob_[1])


@pytest.mark.parametrize("n_lanes", [0, 1])
def test_philox_explicit_matches_oracle(orc, n_lanes):
    """Explicit keywords, free-running: the flattened kernel (n_lanes 0: a warp spreads the auctions and the
    phantom slots of a batch of units over its lanes) and one thread per unit (n_lanes 1) against the oracle;
    shapes that give batches of 1, 2 and 32 units per warp, budgets that bind (exact serial kernel)."""
    rng = np.random.default_rng(3)
    for (K, E, budget) in [(1, 16, 1000.0), (10, 32, 1000.0), (10, 32, 15.0), (7, 3000, 1000.0)]:
        table = make_explicit_table(rng, K)
        env = _env(table, E, seed=77, budget=budget, obs_dtype=torch.float64, n_lanes=n_lanes)
        ob = _oracle_batch(orc, table, E, seed=77, budget=budget)
        for s in range(4):
            bids = np.round(rng.uniform(0.01, 3.0, (E, K)), 2)
            obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
            ref = ob.step(bids, n_threads=4)
            _compare(obs, reward, term, trunc, ref, env, RTOL64)


def test_philox_explicit_large_volume_and_drift(orc):
    """A unit beyond the flattened explicit kernel's volume cap sends its env to the exact serial kernel;
    non-stationary explicit keywords drift from the same ST_UNIT draw in both."""
    rng = np.random.default_rng(13)
    K, E = 4, 6
    table = make_explicit_table(rng, K, E=E)
    table.vol_mean[::2, 1] = 70000.0
    table.vol_std[::2, 1] = 5.0
    mask = [True, False, True, True]
    env = _env(table, E, seed=5, budget=1e9, obs_dtype=torch.float64, updater_mask=mask)
    ob = _oracle_batch(orc, table, E, seed=5, budget=1e9, mask=np.array(mask))
    for s in range(3):
        bids = np.round(rng.uniform(0.01, 3.0, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
    assert int(obs["impressions"].max()) > 1000


def test_drift_matches_oracle(orc):
    rng = np.random.default_rng(21)
    K, E = 16, 24
    table = make_implicit_table(rng, K, 64, E=E)
    mask = rng.random(K) < 0.7
    env = _env(table, E, seed=31, budget=1e5, updater_mask=list(mask), obs_dtype=torch.float64)
    ob = _oracle_batch(orc, table, E, seed=31, budget=1e5, mask=mask)
    for s in range(6):
        bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
        cur = env.keyword_params()
        for n in ("vol_mean", "ctr", "cvr"):
            assert np.array_equal(cur[n], ob.p[n]), n


def test_f32_outputs_and_f32_bids(orc):
    rng = np.random.default_rng(5)
    K, E = 25, 32
    table = make_implicit_table(rng, K, 128)
    env = _env(table, E, seed=1, budget=1e5, obs_dtype=torch.float32)
    ob = _oracle_batch(orc, table, E, seed=1, budget=1e5)
    bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
    obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids.astype(np.float32)).cuda()})
    ref = ob.step(bids, n_threads=4)
    _compare(obs, reward, term, trunc, ref, env, RTOL32)
    assert obs["cost"].dtype == torch.float32


def test_sharding_invariance(orc):
    """Global env ids key the draws: two shards of 20 envs equal one env batch of 40."""
    rng = np.random.default_rng(9)
    K, E = 10, 40
    table = make_implicit_table(rng, K, 64)
    bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
    full = _env(table, E, seed=8, budget=1e5)
    lo = _env(table, E // 2, seed=8, budget=1e5, env_base=0)
    hi = _env(table, E // 2, seed=8, budget=1e5, env_base=E // 2)
    of = full.step({"keyword_bids": torch.from_numpy(bids).cuda()})[0]
    ol = lo.step({"keyword_bids": torch.from_numpy(bids[: E // 2]).cuda()})[0]
    oh = hi.step({"keyword_bids": torch.from_numpy(bids[E // 2:]).cuda()})[0]
    for k in of:
        assert torch.equal(of[k], torch.cat([ol[k], oh[k]])), k


def test_work_list_static_equals_dynamic(orc):
    """The hot kernel's work list: batches pulled from the atomic counter (adc_scratch.work_counter,
    one batch per pull on small steps, chunks of several batches on large ones) or dealt statically
    (no counter: full rounds + small tail batches) -- identical results, step after step, and the
    small case equals the oracle.  Shapes are not multiples of the batch size."""
    rng = np.random.default_rng(21)
    for K, E, steps in ((7, 45, 3), (301, 8200, 2)):   # 315 units; 2.47 M units (chunked pulls)
        table = make_implicit_table(rng, K, 40)
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2)).cuda()
        dyn = _env(table, E, seed=5, budget=1e9)
        sta = _env(table, E, seed=5, budget=1e9, dynamic_work=False)
        ob = _oracle_batch(orc, table, E, seed=5, budget=1e9) if E < 100 else None
        for _ in range(steps):
            od, rd, td, ud, _i = dyn.step({"keyword_bids": bids})
            os_, rs, ts, us, _i = sta.step({"keyword_bids": bids})
            for k in od:
                assert torch.equal(od[k], os_[k]), (k, K, E)
            assert torch.equal(rd, rs) and torch.equal(td, ts) and torch.equal(ud, us)
            if ob is not None:
                _compare(od, rd, td, ud, ob.step(bids.cpu().numpy(), n_threads=2), dyn, RTOL32)
        assert int(od["impressions"].sum()) > 0


def test_bids_beyond_fast_kernel_caps_take_the_exact_route(orc):
    """Bids above 65535 cents overflow the fast kernel's 32-bit lane accumulators by design: such
    envs are routed to the exact serial kernel and still match the oracle."""
    rng = np.random.default_rng(13)
    K, E = 9, 20
    table = make_implicit_table(rng, K, 80)
    env = _env(table, E, seed=3, budget=1e9, obs_dtype=torch.float64)
    ob = _oracle_batch(orc, table, E, seed=3, budget=1e9)
    bids = np.round(rng.uniform(0.2, 1.5, (E, K)), 2)
    bids[::3, 2] = 700.0
    bids[1::4, 5] = 1.0e6
    obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
    ref = ob.step(bids, n_threads=4)
    _compare(obs, reward, term, trunc, ref, env, RTOL64)


@pytest.mark.parametrize("vol,tight_ws", [(900, False), (420, False), (420, True), (200, True)])
@pytest.mark.parametrize("alias", [False, True])
def test_serial_warp_kernel_buffer_overflow_and_zero_budget(orc, alias, vol, tight_ws):
    """What the warp-serial kernel's slab cannot describe takes the direct re-walk: days of more than
    512 auctions (vol 900), sub-steps with more than 61 impressions, chunks of 32 keywords with more
    than 4096 clicked slots in the day (vol 420 with the smallest workspace, `tight_ws`: whole units
    direct, mixed with slab units; with the default workspace the pools grow to 512 slots per keyword and
    nothing overflows); budgets of 0 stop after the first lane of the day (bsim:230-233)."""
    rng = np.random.default_rng(41)
    K, E = 37, 24
    table = make_implicit_table(rng, K, vol)
    table.ctr[:] = rng.uniform(0.7, 1.0, K)
    if vol == 200:
        table.ctr[::2] = rng.uniform(0.05, 0.3, (K + 1) // 2)
    extra = {}
    if tight_ws:
        from adcraft_b200 import _capi
        extra["serial_ws_bytes"] = E * int(_capi.load().adc_serial_slab_bytes(K))
    env = _env(table, E, seed=12, budget=1000.0, budget_alias=alias, obs_dtype=torch.float64, **extra)
    ob = _oracle_batch(orc, table, E, seed=12, budget=1000.0, alias=alias)
    budgets = rng.choice([0.0, 5.0, 60.0, 400.0, 2500.0], size=E)
    for s in range(2):
        bids = np.round(rng.uniform(0.6, 1.6, (E, K)), 2)
        ob.budget[:] = budgets
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda(),
                                                "budget": torch.from_numpy(budgets).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
        np.testing.assert_allclose(env._out["remaining_budget"].cpu().numpy(),
                                   [0.0] * 0 + list(env._out["remaining_budget"].cpu().numpy()))


@pytest.mark.parametrize("budget", [1e5, 3.0])
def test_philox_multi_bidder_matches_oracle(orc, budget):
    """ADC_IMPLICIT_MULTI (the class-default ImplicitKeyword): per-lane Binomial bidder counts, signed
    un-rounded Laplace bids, zero padding below three bidders, negative costs -- CUDA == C oracle."""
    from adcraft_b200 import keywords as kwm
    rng = np.random.default_rng(21)
    K, E = 9, 40
    table = kwm.KeywordTable(
        kwm.IMPLICIT_MULTI, np.full(K, 30.0), np.full(K, 5.0), rng.uniform(-0.1, 0.2, K), rng.uniform(0.05, 0.2, K),
        rng.uniform(0.2, 0.9, K), rng.uniform(0.2, 0.9, K), rng.uniform(0.3, 1.5, K), rng.uniform(0.05, 0.3, K),
        max_bidders=rng.choice([0.0, 1.0, 2.0, 3.0, 5.0, 30.0], K), participation=rng.uniform(0.2, 0.9, K))
    env = _env(table, E, seed=9, budget=budget, obs_dtype=torch.float64, max_days=3)
    params = {n: getattr(table, n) for n in kwm.PARAM_NAMES + ("max_bidders", "participation")}
    ob = orc.BatchOracle(table.kind, E, K, params, seed=9, budget=budget, max_days=3)
    for s in range(4):
        bids = np.round(rng.uniform(0.01, 0.5, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
        ref = ob.step(bids, n_threads=4)
        _compare(obs, reward, term, trunc, ref, env, RTOL64)
    assert int(obs["impressions"].sum()) > 0  # (negative clearing prices are in the rec_multi_* goldens)


@pytest.mark.parametrize("vol", [128, 40])
def test_spread_outcomes_changes_nothing(vol):
    """adc_step_args.spread_outcomes picks the hot kernel's variant that may spread a batch's (unit, group)
    pairs over the lanes: same Philox calls, same masks -- the step is bit-identical either way, on dense
    uneven days (where batches do spread) and on short ones (where none does)."""
    rng = np.random.default_rng(vol)
    K, E = 100, 96
    table = make_implicit_table(rng, K, vol)
    a = _env(table, E, seed=5, budget=1e7, spread_outcomes=True)
    b = _env(table, E, seed=5, budget=1e7, spread_outcomes=False)
    c = _env(table, E, seed=5, budget=1e7)                     # decided from the table
    for s in range(3):
        bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2)).cuda()
        oa, ob_, oc = (x.step({"keyword_bids": bids}) for x in (a, b, c))
        for k in oa[0]:
            assert torch.equal(oa[0][k], ob_[0][k]) and torch.equal(oa[0][k], oc[0][k]), k
        assert torch.equal(oa[1], ob_[1]) and torch.equal(oa[1], oc[1])
    assert c._spread_outcomes() == (1 if vol >= 96 else 0)
    assert int(oa[0]["impressions"].sum()) > 0
