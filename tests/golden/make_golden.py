"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Each fixture holds, for a short episode: the keyword parameters, the canonical bids, the
budget (and whether it was passed as an ndarray -> aliasing double charge), one replay tape per
step, and everything the reference's ``BiddingSimulation.step`` produced for it (observations,
reward, flags, per-lane counts, drifted parameters).  Two families:

* ``rec_*``   the reference ran on its own numpy Generator (+ numpy stand-ins for the three
              unseedable Rust draws); the tape is what it consumed.
* ``phx_*``   the tape was produced by the oracle's Philox "tape function" and replayed
              through the reference with TapeRNG front-ends.

``notebook_lane`` restates the printed outputs of
``adcraft/appendix_bidding_outcomes_example/manual_bidding_example.ipynb`` (cell 2): volume 17
(printed; it came from the unseeded Rust RNG), bid 0.75, keyword 0 right after ``reset(seed=0)``.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
import golden_io  # noqa: E402


def record_case(name, env, bids_seq, budget, steps_meta):
    steps = []
    for bids in bids_seq:
        action = {"keyword_bids": bids}
        if budget is not None:
            action["budget"] = budget
        steps.append(rd.record_step(env, action))
    golden_io.save_case(os.path.join(HERE, name + ".npz"), steps, steps_meta)
    print("wrote", name, "steps", len(steps), "lanes", [s["lanes_run"] for s in steps])


def main():
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    envmod = ref["env"]

    def implicit_env(vol, cvr, K, seed, mask=None, max_days=60):
        env = envmod.bidding_sim_creator(dict(
            keyword_config=rh.experiment_keyword_config(vol, cvr, tmp), num_keywords=K,
            max_days=max_days, updater_mask=mask))
        env.reset(seed=seed)
        return env

    rng = np.random.default_rng(2024)
    bids_for = lambda K, n, hi=1.5: [np.round(rng.uniform(0.05, hi, K), 2) for _ in range(n)]

    # --- reference on its own RNG -----------------------------------------------------------
    record_case("rec_implicit_sparse_k2", implicit_env(16, 0.5, 2, 0), bids_for(2, 4), 1000.0,
                dict(kind="implicit", note="experiment config vol 16 cvr 0.5, reset(seed=0), scalar budget"))
    record_case("rec_implicit_dense_drift_k7", implicit_env(128, 0.8, 7, 1, mask=[True] * 7, max_days=3),
                bids_for(7, 4), np.array([100000.0]),
                dict(kind="implicit", mask=[1] * 7, max_days=3,
                     note="dense, drift on all keywords, ndarray budget, terminates at day 3"))
    record_case("rec_implicit_budget_scalar_k5", implicit_env(64, 0.1, 5, 2), bids_for(5, 3), 30.0,
                dict(kind="implicit", note="binding scalar budget 30"))
    record_case("rec_implicit_budget_alias_k5", implicit_env(64, 0.1, 5, 2), bids_for(5, 3), np.array([30.0]),
                dict(kind="implicit", note="binding ndarray budget 30: double charge + early break"))
    record_case("rec_implicit_partial_mask_k6", implicit_env(64, 0.8, 6, 3, mask=[True, False, True, True, False, False]),
                bids_for(6, 3), 100000.0,
                dict(kind="implicit", mask=[1, 0, 1, 1, 0, 0],
                     note="partial drift mask: zip truncation to num_updates=3"))
    e = envmod.BiddingSimulation(num_keywords=10)
    e.reset(seed=1)
    record_case("rec_explicit_default_k10", e, bids_for(10, 3, 3.0), np.array([1000.0]),
                dict(kind="explicit", note="default env (ExplicitKeyword), reset(seed=1), ndarray budget"))
    e = envmod.BiddingSimulation(num_keywords=1)
    e.reset(seed=0)
    record_case("rec_explicit_default_k1", e, bids_for(1, 4, 3.0), 1000.0,
                dict(kind="explicit", note="single keyword default env, scalar budget"))
    e = envmod.BiddingSimulation(num_keywords=4)
    e.reset(seed=3)
    record_case("rec_explicit_budget_k4", e, bids_for(4, 3, 3.0), 6.0,
                dict(kind="explicit", note="binding scalar budget 6.0"))

    # float32 bids (what the Box(dtype=float32) action space yields): numpy >= 2 keeps them float32
    # through np.maximum / round (env:215), so a bid whose float32 value lies above its cent value
    # WINS ties against the float64 competitor bids.  Own rng: the cases below keep
    # their bids.
    rng32 = np.random.default_rng(32)
    env32 = implicit_env(128, 0.5, 5, 7)
    bids32 = [np.round(rng32.uniform(0.3, 0.9, 5), 2).astype(np.float32) for _ in range(4)]
    record_case("rec_implicit_f32_bids_k5", env32, bids32, 1000.0,
                dict(kind="implicit", f32_bids=True, note="float32 bids: numpy >= 2 tie rule"))
    case32 = golden_io.load_case(os.path.join(HERE, "rec_implicit_f32_bids_k5.npz"))
    differs = 0
    for s in case32.steps:  # the fixture must exercise the rule: float64 semantics give other counts
        kw = orc.KeywordSet(orc.IMPLICIT, *[s.kw_before[n] for n in golden_io.PARAMS])
        t = s.tape
        o = orc.step_replay(kw, s.bid_cents, s.budget, orc.Tape(
            t.volume, t.comp_off, t.comp_cents, t.click_off, np.r_[t.u_click, np.ones(256)], t.conv_off,
            np.r_[t.u_conv, np.ones(256)], t.rev_off, np.r_[t.rev_cents, np.ones(256, np.int32)]))
        differs += int(not np.array_equal(o["impressions"], s.impressions))
    assert differs > 0, "no tie was won in the float32 fixture; change its seed"

    # --- Philox tapes through the reference --------------------------------------------------
    def philox_case(name, kw, budget, alias, mask, n_steps, seed, env_id, hi):
        env = rd.build_replay_env(kw, budget=budget, drift_mask=mask, max_days=3)
        kwc = kw.copy()
        steps = []
        for step in range(n_steps):
            bids = np.round(rng.uniform(0.05, hi, kw.K), 2)
            bc = np.rint(bids * 100).astype(np.int32)
            o = orc.step_philox(kwc, bc, budget, seed=seed, env_id=env_id, step=step,
                                record_cap=8192, budget_alias=alias)
            tape = o["tape"]
            if mask is not None:
                tape.drift = orc.drift_philox(kw.K, seed, env_id, step)
            kw_before = kwc.copy()
            r = rd.replay_step(env, bids, np.array([budget]) if alias else None, tape)
            # the reference agrees with the oracle on this tape, or the fixture is not written
            for f in ("impressions", "clicks", "conversions"):
                assert np.array_equal(np.asarray(r[f], np.int64), np.asarray(o[f], np.int64)), (name, f)
            assert r["reward"] == o["reward"], name
            if mask is not None:
                orc.drift_apply(kwc, np.asarray(mask), tape.drift, kw.vol_std)
            r.update(tape=tape, kw_before=kw_before, bid_cents=bc, budget=float(budget),
                     budget_alias=bool(alias), philox=dict(seed=seed, env_id=env_id, step=step))
            steps.append(r)
        golden_io.save_case(os.path.join(HERE, name + ".npz"), steps,
                            dict(kind="implicit" if kw.kind == orc.IMPLICIT else "explicit",
                                 note="Philox tape function replayed through the reference",
                                 seed=seed, env_id=env_id, max_days=3,
                                 mask=None if mask is None else [int(m) for m in mask]))
        print("wrote", name, "lanes", [s["lanes_run"] for s in steps])

    K = 9
    loc = rng.uniform(0.3, 1.0, K)
    kw_i = orc.KeywordSet(orc.IMPLICIT, np.full(K, 128.0), np.floor(1 + rng.random(K) * 64), loc,
                          np.maximum(0.01, rng.uniform(0.01, 0.3, K) * loc), rng.uniform(0.1, 0.9, K),
                          rng.uniform(0.1, 0.9, K), rng.uniform(0.3, 1.5, K), rng.uniform(0.01, 0.3, K))
    philox_case("phx_implicit_dense_k9", kw_i, 1e5, False, [True] * K, 4, 0x5EED, 17, 1.5)
    philox_case("phx_implicit_budget_alias_k9", kw_i, 12.0, True, None, 3, 0x5EED, 4000000000, 1.5)
    K = 6
    vm = np.floor(rng.uniform(14, 30, K))
    mr = rng.beta(2, 5, K) * 1.5
    kw_e = orc.KeywordSet(orc.EXPLICIT, vm, rng.random(K) * 0.5 * (vm + 1), rng.random(K) * 1.5,
                          rng.beta(5, 5, K) * 25, rng.beta(2, 5, K), rng.beta(5, 2, K), mr,
                          rng.beta(2, 5, K) * mr)
    philox_case("phx_explicit_k6", kw_e, 1000.0, False, [True] * K, 3, 99, 3, 3.0)
    philox_case("phx_explicit_budget_k6", kw_e, 9.0, True, None, 3, 99, 5, 3.0)

    # --- the class-default ImplicitKeyword: m ~ Binomial bidders per lane, signed Laplace bids ----
    rngm = np.random.default_rng(77)
    K = 5
    kw_m = orc.KeywordSet(orc.IMPLICIT_MULTI, np.full(K, 40.0), np.full(K, 6.0), np.array([0.0, 0.05, -0.1, 0.2, 0.0]),
                          np.array([0.1, 0.2, 0.05, 0.1, 0.1]), rngm.uniform(0.2, 0.9, K), rngm.uniform(0.2, 0.9, K),
                          rngm.uniform(0.3, 1.5, K), rngm.uniform(0.05, 0.3, K),
                          max_bidders=np.array([30.0, 2.0, 5.0, 0.0, 30.0]),      # classes:659-662 default 30; 2 and 0:
                          participation=np.array([0.6, 0.5, 0.3, 0.6, 0.6]))      # zero-padded short auctions
    envm = rd.build_multi_env(kw_m, seed=5, budget=1000.0)
    record_case("rec_multi_default_k5", envm, [np.round(rngm.uniform(0.01, 0.5, K), 2) for _ in range(3)], 1000.0,
                dict(kind="multi", note="class-default ImplicitKeyword: Binomial bidders per lane, signed Laplace, "
                                        "zero padding for m < 3, negative costs (helpers:116-180)"))
    envm = rd.build_multi_env(kw_m, seed=6, budget=1000.0)
    record_case("rec_multi_budget_k5", envm, [np.round(rngm.uniform(0.05, 0.5, K), 2) for _ in range(3)], np.array([2.5]),
                dict(kind="multi", note="binding ndarray budget: double charge + early break"))

    def philox_multi(name, budget, alias, n_steps, seed, env_id):
        env = rd.build_replay_env(kw_m, budget=budget, max_days=3)
        steps = []
        for step in range(n_steps):
            bids = np.round(rngm.uniform(0.02, 0.5, K), 2)
            bc = np.rint(bids * 100).astype(np.int32)
            o = orc.step_philox(kw_m, bc, budget, seed=seed, env_id=env_id, step=step, record_cap=8192, budget_alias=alias)
            r = rd.replay_step(env, bids, np.array([budget]) if alias else None, o["tape"])
            for f in ("impressions", "clicks", "conversions"):
                assert np.array_equal(np.asarray(r[f], np.int64), np.asarray(o[f], np.int64)), (name, f)
            assert r["reward"] == o["reward"], name
            r.update(tape=o["tape"], kw_before=kw_m.copy(), bid_cents=bc, budget=float(budget), budget_alias=bool(alias))
            steps.append(r)
        golden_io.save_case(os.path.join(HERE, name + ".npz"), steps,
                            dict(kind="multi", note="Philox tape function replayed through the reference",
                                 seed=seed, env_id=env_id, max_days=3, mask=None))
        print("wrote", name, "lanes", [s["lanes_run"] for s in steps])

    philox_multi("phx_multi_k5", 1000.0, False, 3, 21, 9)
    philox_multi("phx_multi_budget_k5", 1.5, True, 3, 21, 10)

    # --- ideal-profit estimator of the AKNCP / NCP metrics (experiment_metrics.py:20-61) -------
    met = ref["metrics"]
    env = implicit_env(128, 0.8, 7, 1, mask=[True] * 7)
    for _ in range(3):  # drifted parameters: the estimator reads the current ones
        env.step({"keyword_bids": np.full(7, 0.75), "budget": 100000.0})
    allowed_bids = np.arange(0.01, 3.00, 0.01)  # run_heatmap_experiments.ipynb cell 3
    kwset = rd.keywordset_from_env(env)
    samples, irs, cpcs, best, frac, arg = [], [], [], [], [], []
    for k, (kwd, params) in enumerate(zip(env.keywords, env.keyword_params)):
        env.np_random.log = log = []
        ir, cpc = met.get_implicit_kw_bid_cpc_impressions(kwd, allowed_bids)
        env.np_random.log = None
        assert len(log) == 1 and log[0][0] == "laplace"
        c = np.around(np.maximum(np.abs(log[0][2]), 0.0).astype(float), 2).ravel()  # helpers:108-113
        samples.append(np.rint(c * 100).astype(np.int32))
        b, f, a = met.get_max_expected_bid_profits(params, cpc, ir)
        irs.append(ir); cpcs.append(cpc); best.append(b); frac.append(f); arg.append(a)
    np.savez_compressed(os.path.join(HERE, "ideal_profit.npz"), allowed_bids=allowed_bids,
                        samples_cents=np.stack(samples), impression_rate=np.stack(irs), expected_cpc=np.stack(cpcs),
                        ideal_profit=np.array(best), positive_frac=np.array(frac), best_bid_index=np.array(arg),
                        **{"kw_" + n: getattr(kwset, n) for n in golden_io.PARAMS})
    print("wrote ideal_profit", np.round(best, 3))

    # --- notebook known-answer lane ---------------------------------------------------------
    env = implicit_env(16, 0.5, 2, 0, max_days=10)
    kw0 = env.keywords[0]
    env.np_random.log = log = []
    comp = kw0.bid_distribution(1, 17)
    wins = comp[comp <= 0.75]
    clicks = kw0.rng.random((len(wins))) <= kw0.buyside_ctr
    convs = env.keywords[0].rng.random((clicks.sum())) <= kw0.sellside_paid_ctr
    revs = kw0.reward_distribution_sampler(convs.sum())
    env.np_random.log = None
    # printed by the notebook (manual_bidding_example.ipynb:89-125)
    printed_comp = [0.67, 0.6, 0.62, 0.81, 0.56, 0.68, 0.46, 0.76, 0.74, 0.57, 0.79, 0.42, 0.6, 0.52,
                    0.63, 0.74, 0.56]
    assert np.allclose(comp.ravel(), printed_comp)
    assert len(wins) == 14 and clicks.sum() == 12 and convs.sum() == 5
    assert np.allclose(revs, [0.88, 0.82, 1.43, 1.41, 1.64])
    assert abs(wins[clicks].sum() - 7.209999999999999) < 1e-12
    np.savez(os.path.join(HERE, "notebook_lane.npz"),
             comp_cents=np.rint(comp.ravel() * 100).astype(np.int32),
             u_click=np.asarray(log[1][2]), u_conv=np.asarray(log[2][2]),
             rev_cents=np.rint(revs * 100).astype(np.int32),
             ctr=kw0.buyside_ctr, cvr=kw0.sellside_paid_ctr, bid_cents=75, volume=17,
             impressions=14, clicks=12, conversions=5, cost=7.209999999999999, revenue=6.18,
             keyword_params=np.array(env.reset()[1]["keyword_params"]))
    print("wrote notebook_lane")


if __name__ == "__main__":
    main()
