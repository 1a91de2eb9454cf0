"""GPU parity against the committed reference goldens, through the C ABI.

* every fixture is replayed through adc_step_replay (tape mode) and must reproduce what the
  unmodified reference produced on that tape: integers bit-exact, floats to 1e-6 relative;
* the phx_* fixtures are additionally run FREE-RUNNING (adc_step_philox with the fixture's
  seed / env id / step): the CUDA Philox path must land on the reference's outputs too.
"""
import glob
import os

import numpy as np
import pytest

import golden_io

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [c for c in sorted(glob.glob(os.path.join(HERE, "golden", "[rp][eh][cx]_*.npz")))
         if not c.endswith("notebook_lane.npz")]
IDS = [os.path.basename(c)[:-4] for c in CASES]


def _table(case, step, E):
    from adcraft_b200 import keywords as kwm
    cols = [np.tile(step.kw_before[n], (E, 1)) for n in golden_io.PARAMS]
    extra = {n: np.tile(v, (E, 1)) for n, v in step.kw_extra.items()}
    return kwm.KeywordTable(case.kind, *cols, **extra)


def _make_env(case, E, **kw):
    from adcraft_b200.vector_env import VectorBiddingSimulation
    s0 = case.steps[0]
    mask = case.meta.get("mask")
    env = VectorBiddingSimulation(
        E, num_keywords=case.K, keywords=_table(case, s0, E), budget=s0.budget,
        max_days=case.meta.get("max_days", 60), updater_mask=None if mask is None else [bool(m) for m in mask],
        budget_alias=bool(s0.budget_alias), obs_dtype=torch.float64, autoreset=False, device="cuda",
        f32_ties=bool(case.meta.get("f32_bids")), **kw)
    env.reset()
    return env


def _check(case, s, obs, reward, term, trunc, env, E, cum):
    for e in range(E):
        for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                     ("sellside_conversions", "conversions")):
            assert np.array_equal(obs[a][e].cpu().numpy(), getattr(s, b).astype(np.int32)), (a, e)
        np.testing.assert_allclose(obs["cost"][e].cpu().numpy(), s.cost, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(obs["revenue"][e].cpu().numpy(), s.revenue, rtol=1e-6, atol=1e-9)
        scale = 1.0 + np.abs(s.cost).sum() + np.abs(s.revenue).sum()
        assert abs(float(reward[e]) - s.reward) <= 1e-6 * scale
        assert abs(float(obs["cumulative_profit"][e, 0]) - s.cumulative_profit) <= 1e-6 * (scale + abs(cum))
        assert int(obs["days_passed"][e, 0]) == int(s.days_passed)
        assert bool(term[e]) == bool(s.terminated) and bool(trunc[e]) == bool(s.truncated)
    if s.kw_after is not None and case.meta.get("mask") is not None:
        cur = env.keyword_params()
        for n in ("vol_mean", "ctr", "cvr"):
            for e in range(E):
                assert np.array_equal(cur[n][e], s.kw_after[n]), n


@pytest.mark.parametrize("E", [1, 3])
@pytest.mark.parametrize("force_serial", [False, True, "packed"])
@pytest.mark.parametrize("path", CASES, ids=IDS)
def test_replay_reproduces_reference(path, E, force_serial):
    from adcraft_b200.tape import DeviceTape
    case = golden_io.load_case(path)
    packed = force_serial == "packed"
    if packed:
        force_serial = False
        if case.meta["kind"] != "implicit":
            pytest.skip("packed records exist for implicit keywords only")
    env = _make_env(case, E)
    cum = 0.0
    for s in case.steps:
        tape = DeviceTape.from_host([s.tape] * E, "cuda", pack=packed)
        bids = torch.from_numpy(np.tile(s.bid_cents / 100.0, (E, 1))).cuda()
        if case.meta.get("f32_bids"):  # the tie rule needs the bids in float32 (adc_step_args.f32_ties)
            bids = torch.from_numpy(np.tile(s.bid_cents.astype(np.float32) / np.float32(100), (E, 1))).cuda()
        budget = torch.full((E,), s.budget, dtype=bids.dtype, device="cuda")  # passed every step
        obs, reward, term, trunc, _ = env.step_replay({"keyword_bids": bids, "budget": budget}, tape,
                                                      force_serial=force_serial)
        _check(case, s, obs, reward, term, trunc, env, E, cum)
        cum = s.cumulative_profit


PHX = [c for c in CASES if os.path.basename(c).startswith("phx_")]


@pytest.mark.parametrize("n_lanes", [0, 8])
@pytest.mark.parametrize("path", PHX, ids=[os.path.basename(c)[:-4] for c in PHX])
def test_free_running_philox_reproduces_reference(path, n_lanes):
    case = golden_io.load_case(path)
    env = _make_env(case, 1, seed=case.meta["seed"], env_base=case.meta["env_id"], n_lanes=n_lanes)
    cum = 0.0
    for s in case.steps:
        bids = torch.from_numpy((s.bid_cents / 100.0)[None]).cuda()
        budget = torch.full((1,), s.budget, dtype=torch.float64, device="cuda")
        obs, reward, term, trunc, _ = env.step({"keyword_bids": bids, "budget": budget})
        _check(case, s, obs, reward, term, trunc, env, 1, cum)
        cum = s.cumulative_profit
