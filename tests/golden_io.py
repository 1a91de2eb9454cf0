"""(De)serialisation of golden episodes (tests/golden/*.npz).  Test infrastructure."""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import Dict, List

import numpy as np

PARAMS = ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std")
TAPE_FIELDS = ("volume", "comp_off", "comp_cents", "click_off", "u_click", "conv_off", "u_conv",
               "rev_off", "rev_cents", "impr", "cost_off", "cost", "drift", "comp_f64")
EXTRA_PARAMS = ("max_bidders", "participation")  # multi-bidder (class-default) implicit keywords
KIND_IDS = {"implicit": 0, "explicit": 1, "multi": 2}
OUT_FIELDS = ("impressions", "clicks", "conversions", "cost", "revenue", "profit", "lane_I", "lane_B",
              "lane_S")
SCALARS = ("reward", "cumulative_profit", "days_passed", "terminated", "truncated", "lanes_run",
           "budget", "budget_alias")


def save_case(path: str, steps: List[Dict], meta: Dict) -> None:
    arrs = {}
    for i, s in enumerate(steps):
        p = f"s{i}_"
        for n in PARAMS:
            arrs[p + "kwb_" + n] = np.asarray(getattr(s["kw_before"], n), np.float64)
            if "kw_after" in s:
                arrs[p + "kwa_" + n] = np.asarray(getattr(s["kw_after"], n), np.float64)
        for n in EXTRA_PARAMS:
            if getattr(s["kw_before"], n, None) is not None:
                arrs[p + "kwb_" + n] = np.asarray(getattr(s["kw_before"], n), np.float64)
        arrs[p + "bid_cents"] = np.asarray(s["bid_cents"], np.int32)
        for n in TAPE_FIELDS:
            v = getattr(s["tape"], n, None)
            if v is not None:
                arrs[p + "tape_" + n] = np.asarray(v)
        for n in OUT_FIELDS:
            arrs[p + n] = np.asarray(s[n])
        arrs[p + "scalars"] = np.array([float(s[n]) for n in SCALARS], np.float64)
        if "info_outcomes" in s:  # info["bidding_outcomes"], the reference's string (lib.rs:250-275)
            arrs[p + "info_outcomes"] = np.array(str(s["info_outcomes"]))
    meta = dict(meta, n_steps=len(steps), kind_id=KIND_IDS[meta["kind"]])
    arrs["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(path, **arrs)


def load_case(path: str) -> SimpleNamespace:
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    steps = []
    for i in range(meta["n_steps"]):
        p = f"s{i}_"
        s = SimpleNamespace()
        s.kw_before = {n: z[p + "kwb_" + n] for n in PARAMS}
        s.kw_extra = {n: z[p + "kwb_" + n] for n in EXTRA_PARAMS if p + "kwb_" + n in z}
        s.kw_after = {n: z[p + "kwa_" + n] for n in PARAMS} if p + "kwa_vol_mean" in z else None
        s.bid_cents = z[p + "bid_cents"]
        s.tape = SimpleNamespace(**{n: (z[p + "tape_" + n] if p + "tape_" + n in z else None)
                                    for n in TAPE_FIELDS})
        for n in OUT_FIELDS:
            setattr(s, n, z[p + n])
        for n, v in zip(SCALARS, z[p + "scalars"]):
            setattr(s, n, float(v))
        s.info_outcomes = str(z[p + "info_outcomes"]) if p + "info_outcomes" in z else None
        steps.append(s)
    return SimpleNamespace(meta=meta, kind=meta["kind_id"], steps=steps, K=len(steps[0].bid_cents))
