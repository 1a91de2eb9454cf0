"""TEST INFRASTRUCTURE ONLY -- drive the unmodified reference env in record / replay mode.

record:  run ``BiddingSimulation.step`` on its own numpy Generator (plus the rust shim's
         numpy stand-ins for the unseedable Rust draws) and turn the logged draws into a
         consumption-ordered :class:`oracle.oracle.Tape`.
replay:  build reference keywords whose Generators are :class:`TapeRNG` objects and run the
         same unmodified ``step`` on a given tape.

Per-lane outcomes are observed by wrapping ``bidding_simulation.simulate_epoch_of_bidding``
(the wrapper only copies its return value).
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import numpy as np

from . import ref_harness as rh
from .oracle import EXPLICIT, IMPLICIT, IMPLICIT_MULTI, SUBSTEPS, KeywordSet, Tape


def _cents(x) -> np.ndarray:
    return np.rint(np.asarray(x, dtype=np.float64) * 100.0).astype(np.int64)


class _LaneSpy:
    """Context manager collecting every lane outcome of one reference step."""

    def __init__(self):
        self.lanes: List[dict] = []

    def __enter__(self):
        self.bsim = rh.load_reference()["bsim"]
        self.orig = self.bsim.simulate_epoch_of_bidding

        def spy(*a, **k):
            out = self.orig(*a, **k)
            self.lanes.append(dict(
                impressions=int(out["impressions"]), clicks=int(out["buyside_clicks"]),
                conversions=int(out["sellside_conversions"]),
                costs=np.array(out["costs"], dtype=np.float64),
                revenues=np.array(out["revenues"], dtype=np.float64),
                profit=float(out["profit"])))
            return out

        self.bsim.simulate_epoch_of_bidding = spy
        return self

    def __exit__(self, *exc):
        self.bsim.simulate_epoch_of_bidding = self.orig


def _collect(env, K, obs, reward, term, trunc, info, spy) -> Dict[str, object]:
    lane_I = np.zeros((SUBSTEPS, K), np.int32)
    lane_B = np.zeros((SUBSTEPS, K), np.int32)
    lane_S = np.zeros((SUBSTEPS, K), np.int32)
    profit = np.zeros(K)
    for l, ln in enumerate(spy.lanes):
        t, k = divmod(l, K)
        lane_I[t, k], lane_B[t, k], lane_S[t, k] = ln["impressions"], ln["clicks"], ln["conversions"]
    # combine_outcomes accumulates profit lane by lane starting from 0.0 (bsim:138-139)
    for l, ln in enumerate(spy.lanes):
        profit[l % K] += ln["profit"]
    return dict(
        impressions=np.asarray(obs["impressions"], np.int64), clicks=np.asarray(obs["buyside_clicks"], np.int64),
        conversions=np.asarray(obs["sellside_conversions"], np.int64),
        cost=np.asarray(obs["cost"], np.float64), revenue=np.asarray(obs["revenue"], np.float64),
        cumulative_profit=float(np.asarray(obs["cumulative_profit"]).ravel()[0]),
        days_passed=int(np.asarray(obs["days_passed"]).ravel()[0]),
        reward=float(np.asarray(reward).ravel()[0]), terminated=bool(term), truncated=bool(np.asarray(trunc).ravel()[0]),
        lane_I=lane_I, lane_B=lane_B, lane_S=lane_S, profit=profit, lanes_run=len(spy.lanes),
        info_bids=[float(b) for b in info["bids"]], info_outcomes=info["bidding_outcomes"],
        info_params=info["keyword_params"])


def keywordset_from_env(env) -> KeywordSet:
    """Current (possibly drifted) parameters of a reference env as SoA columns."""
    ref = rh.load_reference()
    if getattr(env, "_adc_multi", None) is not None:
        kw = env._adc_multi.copy()
        kw.ctr = np.array([float(k.buyside_ctr) for k in env.keywords])
        kw.cvr = np.array([float(k.sellside_paid_ctr) for k in env.keywords])
        kw.vol_mean = np.array([float(p[0][0]) for p in env.keyword_params])
        return kw
    explicit = isinstance(env.keywords[0], ref["classes"].ExplicitKeyword)
    kp = env.keyword_params
    cols = dict(
        vol_mean=[float(p[0][0]) for p in kp], vol_std=[float(p[0][1]) for p in kp],
        p1=[float(p[1]) for p in kp],
        p2=[float(p[2]) if explicit else 1.0 / float(p[2]) for p in kp],
        ctr=[float(k.buyside_ctr) for k in env.keywords],
        cvr=[float(k.sellside_paid_ctr) for k in env.keywords],
        rev_mean=[float(p[5]) for p in kp], rev_std=[float(p[6]) for p in kp])
    return KeywordSet(EXPLICIT if explicit else IMPLICIT, **cols)


def record_step(env, action) -> Dict[str, object]:
    """One unmodified reference step on its own RNG; returns outputs + the tape it consumed."""
    ref = rh.load_reference()
    shim = ref["shim"]
    K = env.num_keywords
    explicit = isinstance(env.keywords[0], ref["classes"].ExplicitKeyword)
    multi = getattr(env, "_adc_multi", None) is not None  # built by build_multi_env below
    kw_before = keywordset_from_env(env)
    # env:197-199 -- read BEFORE the step: an ndarray budget is mutated in place by the lanes
    budget_in = action.get("budget", env.budget)
    budget = float(np.asarray(np.round(budget_in, 2), dtype=float).ravel()[0])
    budget_alias = isinstance(budget_in, np.ndarray) and budget_in.ndim >= 1
    log: List[tuple] = []
    env.np_random.log = log
    shim.log = log
    try:
        with _LaneSpy() as spy:
            obs, reward, term, trunc, info = env.step(action)
    finally:
        env.np_random.log = None
        shim.log = None
    out = _collect(env, K, obs, reward, term, trunc, info, spy)

    it = iter(log)
    ev = list(log)
    pos = 0
    volume = []
    for _ in range(K):
        name, _args, val = ev[pos]; pos += 1
        assert name == "rust.volume", name
        volume.append(int(val))
    comp = [[] for _ in range(K)]
    ucl = [[] for _ in range(K)]
    ucv = [[] for _ in range(K)]
    rev = [[] for _ in range(K)]
    cost = [[] for _ in range(K)]
    compf = [[] for _ in range(K)]
    impr = np.zeros((K, SUBSTEPS), np.int32)
    lane = 0
    while pos < len(ev) and ev[pos][0] != "uniform":
        t, k = divmod(lane, K)
        if explicit:
            name, _a, val = ev[pos]; pos += 1
            assert name == "rust.binomial", name
            impr[k, t] = int(val)
            if int(val) >= 1:
                name, _a, val = ev[pos]; pos += 1
                assert name == "rust.cost_create", name
                cost[k].extend(np.asarray(val, np.float64).ravel().tolist())
        elif multi:
            name, _a, val = ev[pos]; pos += 1
            assert name == "binomial", name
            m = int(val)
            impr[k, t] = m  # the lane's bidder count rides in `impr`
            name, _a, val = ev[pos]; pos += 1
            assert name == "laplace", name
            val = np.asarray(val, np.float64)  # [m, n]: signed, un-rounded (classes:683-686)
            compf[k].extend((val.max(axis=0) if m > 0 else np.zeros(val.shape[1])).tolist())
            comp[k].extend([0] * val.shape[1])
        else:
            name, _a, val = ev[pos]; pos += 1
            assert name == "laplace", name
            c = np.around(np.maximum(np.abs(val), 0.0).astype(float), 2)  # helpers:108-113
            comp[k].extend(_cents(c).ravel().tolist())
        name, _a, val = ev[pos]; pos += 1
        assert name == "random", name
        ucl[k].extend(np.asarray(val).ravel().tolist())
        name, _a, val = ev[pos]; pos += 1
        assert name == "random", name
        ucv[k].extend(np.asarray(val).ravel().tolist())
        name, _a, val = ev[pos]; pos += 1
        assert name == "normal", name
        r = np.around(np.maximum(val, 0.01).astype(float), 2)  # helpers:68-70
        rev[k].extend(_cents(r).ravel().tolist())
        lane += 1
    assert lane == out["lanes_run"], (lane, out["lanes_run"])
    drift = None
    if pos < len(ev):
        drift = np.zeros((3, K))
        for c in range(3):
            name, _a, val = ev[pos]; pos += 1
            assert name == "uniform", name
            drift[c, :len(val)] = val
    assert pos == len(ev)
    tape = Tape.from_lists(volume, comp, ucl, ucv, rev,
                           impr=impr if explicit else None, cost=cost if explicit else None, drift=drift)
    if multi:
        tape.impr = impr
        tape.comp_f64 = np.asarray([x for row in compf for x in row], np.float64)
    out["tape"] = tape
    out["kw_before"] = kw_before
    out["kw_after"] = keywordset_from_env(env)
    out["bid_cents"] = _cents(out["info_bids"]).astype(np.int32)
    out["budget"] = budget
    out["budget_alias"] = bool(budget_alias)
    out["budget_after"] = float(np.asarray(env.budget).ravel()[0])
    return out


class _ShimTapeSource:
    """Feeds the rust shim's three random helpers from a Tape (call order of one step)."""

    def __init__(self, tape: Tape, K: int):
        self.tape, self.K = tape, K
        self.i_vol = 0
        self.lane = -1
        self.n_cost = [0] * K

    def volume(self, mean, std):
        v = int(self.tape.volume[self.i_vol]); self.i_vol += 1
        return v

    def binomial(self, n, p):
        self.lane += 1
        t, k = divmod(self.lane, self.K)
        return int(self.tape.impr[k, t])

    def costs(self, x, n):
        _t, k = divmod(self.lane, self.K)
        a = int(self.tape.cost_off[k]) + self.n_cost[k]
        self.n_cost[k] += n
        assert a + n <= int(self.tape.cost_off[k + 1]), "tape exhausted: explicit costs"
        return self.tape.cost[a:a + n]


def multi_keyword(kw: KeywordSet, k: int, rng):
    """The reference's class-default ImplicitKeyword (synthetic_kw_classes.py:578-688): bidders ~
    Binomial(max_bidders, participation) per lane, signed Laplace(bid_loc, bid_scale) bids -- only the
    parameters are given, the two distributions are the class's own defaults."""
    ref = rh.load_reference()
    helpers = ref["helpers"]
    vol = (kw.vol_mean[k], kw.vol_std[k])
    key = ref["classes"].ImplicitKeyword({
        "rng": rng, "max_bidders": int(kw.max_bidders[k]), "participation_rate": float(kw.participation[k]),
        "bid_loc": float(kw.p1[k]), "bid_scale": float(kw.p2[k]),
        "sellside_paid_ctr": float(kw.cvr[k]), "buyside_ctr": float(kw.ctr[k]),
        "volume_sampler": helpers.nonneg_int_normal_sampler(rng, the_mean=vol[0], std=vol[1]),
        "reward_distribution_sampler": helpers.rev_normal(float(kw.rev_mean[k]), std_dev=float(kw.rev_std[k]), rng=rng),
    }, verbose=False)
    return key, (vol, kw.p1[k], kw.p2[k], kw.ctr[k], kw.cvr[k], kw.rev_mean[k], kw.rev_std[k])


def build_multi_env(kw: KeywordSet, seed: int, *, budget=1000.0, max_days=60):
    """A reference BiddingSimulation of class-default ImplicitKeywords on ONE RecordingRNG (record mode)."""
    ref = rh.load_reference()
    env = ref["env"].BiddingSimulation(num_keywords=kw.K, budget=budget, max_days=max_days)
    env._np_random = rh.RecordingRNG(np.random.PCG64(np.random.SeedSequence(seed)))
    kws, params = [], []
    for k in range(kw.K):
        key, p = multi_keyword(kw, k, env._np_random)
        kws.append(key); params.append(list(p))
    env.keywords, env.keyword_params = kws, params
    env._adc_multi = kw.copy()
    env._have_keywords = True
    env.current_day, env.cumulative_profit = 0, 0.0
    return env


def build_replay_env(kw: KeywordSet, *, budget=1000.0, max_days=60, loss_threshold=10000.0,
                     drift_mask=None, drift_mag=(0.03, 0.03, 0.03), cum_profit=0.0, day=0):
    """A reference BiddingSimulation whose keywords draw from TapeRNGs."""
    ref = rh.load_reference()
    utils, envmod = ref["utils"], ref["env"]
    K = kw.K
    env = envmod.BiddingSimulation(
        num_keywords=K, budget=budget, max_days=max_days, loss_threshold=loss_threshold,
        updater_params=[["vol", drift_mag[0]], ["ctr", drift_mag[1]], ["cvr", drift_mag[2]]],
        updater_mask=None if drift_mask is None else [bool(m) for m in drift_mask])
    env._np_random = rh.TapeRNG()
    rngs, kws, params = [], [], []
    for k in range(K):
        r = rh.TapeRNG()
        vol = (kw.vol_mean[k], kw.vol_std[k])
        if kw.kind == IMPLICIT:
            key, p = utils.generate_implicit_keyword_from_params(
                vol, kw.p1[k], kw.p2[k], kw.ctr[k], kw.cvr[k], kw.rev_mean[k], kw.rev_std[k], r)
        elif kw.kind == IMPLICIT_MULTI:
            key, p = multi_keyword(kw, k, r)
        else:
            key, p = utils.generate_keyword_from_params(
                vol, kw.p1[k], kw.p2[k], kw.ctr[k], kw.cvr[k], kw.rev_mean[k], kw.rev_std[k], r)
        rngs.append(r); kws.append(key); params.append(list(p))
    env.keywords, env.keyword_params = kws, params
    if kw.kind == IMPLICIT_MULTI:
        env._adc_multi = kw.copy()
    env._have_keywords = True
    env.current_day = day
    env.cumulative_profit = cum_profit
    env._tape_rngs = rngs
    return env


def replay_step(env, bids_dollars, budget, tape: Tape) -> Dict[str, object]:
    """Run the unmodified reference step of `env` (from build_replay_env) on `tape`."""
    ref = rh.load_reference()
    shim = ref["shim"]
    K = env.num_keywords
    tape.normalise()
    multi = getattr(env, "_adc_multi", None) is not None
    for k, r in enumerate(env._tape_rngs):
        if multi:
            r.bidders = np.asarray(tape.impr[k], np.int64)  # lanes of keyword k run in sub-step order
            r.load(tape.comp_f64[tape.comp_off[k]:tape.comp_off[k + 1]],
                   tape.u_click[tape.click_off[k]:tape.click_off[k + 1]],
                   tape.u_conv[tape.conv_off[k]:tape.conv_off[k + 1]],
                   tape.rev_cents[tape.rev_off[k]:tape.rev_off[k + 1]].astype(np.float64) / 100.0)
            continue
        r.load(tape.comp_cents[tape.comp_off[k]:tape.comp_off[k + 1]].astype(np.float64) / 100.0,
               tape.u_click[tape.click_off[k]:tape.click_off[k + 1]],
               tape.u_conv[tape.conv_off[k]:tape.conv_off[k + 1]],
               tape.rev_cents[tape.rev_off[k]:tape.rev_off[k + 1]].astype(np.float64) / 100.0)
    env._np_random.reset_cursors()
    if tape.drift is not None and env.updater_mask is not None:
        nu = int(np.sum(env.updater_mask))
        env._np_random.drift = [tape.drift[c, :nu] for c in range(3)]
    old_source = shim.source
    shim.source = ("tape", _ShimTapeSource(tape, K))
    action = {"keyword_bids": np.asarray(bids_dollars, dtype=np.float64)}
    if budget is not None:
        action["budget"] = np.array([budget], dtype=np.float64)
    try:
        with _LaneSpy() as spy:
            obs, reward, term, trunc, info = env.step(action)
    finally:
        shim.source = old_source
    out = _collect(env, K, obs, reward, term, trunc, info, spy)
    out["kw_after"] = keywordset_from_env(env)
    return out
