"""Generate tests/golden/baseline_traces.npz from the UNMODIFIED reference baseline bidders
(adcraft/baselines/interpolated_expectations.py:298-515; build container only).

    python tests/golden/make_baseline_traces.py

For each strategy a trace of T steps over K keywords: the observation fed to
``update_all_caches``, the previous bids, the uniforms the reference's Generator consumed inside
``sample_action`` (re-drawn from the saved bit-generator state), and the action it returned.  The
GPU test replays the observations and uniforms through the vectorised policies on the device and
must land on the reference's bids and budgets."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OBS = ("impressions", "buyside_clicks", "cost", "sellside_conversions", "revenue")


def fake_obs(rng, K):
    imp = rng.integers(0, 60, K)
    clicks = np.minimum(imp, rng.integers(0, 25, K)) * (rng.random(K) < 0.7)
    conv = np.minimum(clicks, rng.integers(0, 12, K)) * (rng.random(K) < 0.6)
    rev = np.round(conv * rng.uniform(0.3, 1.5, K), 2)
    cost = np.round(clicks * rng.uniform(0.1, 0.8, K), 2)
    return dict(impressions=imp.astype(float), buyside_clicks=clicks.astype(float), cost=cost,
                sellside_conversions=conv.astype(float), revenue=rev)


def main():
    rh.load_reference()
    ie = importlib.import_module("adcraft.baselines.interpolated_expectations")
    out = {}
    # ---- NaiveZeroMarginStrategy (:442-515)
    rng = np.random.default_rng(0)
    K, T = 12, 40
    ref = ie.NaiveZeroMarginStrategy(K, seed=123)
    bids = np.full(K, 0.01)
    rec = {n: [] for n in OBS + ("prev_bids", "uniforms", "bids", "budget")}
    for _ in range(T):
        obs = fake_obs(rng, K)
        ref.update_all_caches({"keyword_bids": bids}, {k: v.copy() for k, v in obs.items()})
        need = [i for i in range(K) if ref.caches[i]["num_rpc_obs"] < 1]
        state = ref.rng.bit_generator.state
        with np.errstate(divide="ignore"):
            act = ref.sample_action()
        ref.rng.bit_generator.state = state
        u = np.ones(K)
        u[need] = ref.rng.random(len(need))
        for n in OBS:
            rec[n].append(obs[n])
        rec["prev_bids"].append(bids.copy()); rec["uniforms"].append(u)
        rec["bids"].append(np.asarray(act["keyword_bids"], float)); rec["budget"].append(float(act["budget"]))
        bids = np.asarray(act["keyword_bids"], float)
    out.update({"zm_" + n: np.asarray(v) for n, v in rec.items()})
    # ---- NaiveInterpolationStrategy (:298-439)
    rng = np.random.default_rng(1)
    K, T = 9, 60
    ref = ie.NaiveInterpolationStrategy(K, seed=77)
    bids = np.full(K, 0.01)
    rec = {n: [] for n in OBS + ("prev_bids", "uniforms", "bids", "budget")}
    for step in range(T):
        obs = fake_obs(rng, K)
        if step % 7 == 3:
            for n in ("buyside_clicks", "sellside_conversions", "revenue", "cost"):
                obs[n][:] = 0
        ref.update_all_caches({"keyword_bids": bids}, {k: v.copy() for k, v in obs.items()})
        has = []
        for i in range(K):
            m_ref, _ = ref.get_expected_margin_from_cache(i)
            has.append(ref.get_profit_acquisition_function(np.array(m_ref), index=i) is not None)
        state = ref.rng.bit_generator.state
        act = ref.sample_action()
        ref.rng.bit_generator.state = state
        u = np.zeros(K)
        u[np.array(has)] = ref.rng.random(int(np.sum(has)))
        for n in OBS:
            rec[n].append(obs[n])
        rec["prev_bids"].append(bids.copy()); rec["uniforms"].append(u)
        rec["bids"].append(np.asarray(act["keyword_bids"], float)); rec["budget"].append(float(act["budget"]))
        bids = np.asarray(act["keyword_bids"], float)
    out.update({"ni_" + n: np.asarray(v) for n, v in rec.items()})
    np.savez_compressed(os.path.join(HERE, "baseline_traces.npz"), **out)
    print("wrote baseline_traces.npz", {k: v.shape for k, v in out.items() if k.endswith("bids")})


if __name__ == "__main__":
    main()
