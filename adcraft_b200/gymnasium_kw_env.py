"""E = 1 adapter with the reference's single-env interface (numpy in, numpy out).

Mirror of ``adcraft/gymnasium_kw_env.py:22-363``: same constructor keywords, ``reset(*, seed,
options) -> (obs, info)``, ``step(action) -> (obs, reward, terminated, truncated, info)``, same
observation keys / dtypes (int64 counts, float64 money, ``cumulative_profit`` and ``days_passed``
of shape ``(1,)``), same info keys, ``render`` / ``close`` / ``set_updater_mask``, and the factory
``bidding_sim_creator``.  All arithmetic happens on the GPU through VectorBiddingSimulation.

Differences, by construction: draws come from Philox counters (the reference's Rust RNG cannot be
seeded, so its numbers are not reproducible either).  ``info["bidding_outcomes"]`` carries every
click's cost, the per-click revenues and the impression share like the reference's string: the
single-env adapter always takes the exact serial kernel, which records them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from .spaces import HAVE_GYMNASIUM, get_action_space, get_observation_space
from .vector_env import DEFAULT_UPDATER_PARAMS, VectorBiddingSimulation, _sum_list

if HAVE_GYMNASIUM:  # pragma: no cover - gymnasium is not in this image
    import gymnasium as _gym
    _Base = _gym.Env
else:
    _Base = object


class BiddingSimulation(_Base):
    metadata = {"render_modes": ["ansi"]}

    def __init__(self, keyword_config: Optional[Dict] = None, num_keywords: int = 10,
                 budget: float = 1000.0, render_mode: Optional[str] = None,
                 loss_threshold: float = 10000.0, max_days: int = 60,
                 updater_params: List[List] = DEFAULT_UPDATER_PARAMS,
                 updater_mask: Optional[List[bool]] = None, **kwargs) -> None:
        self._vec = VectorBiddingSimulation(
            1, keyword_config=keyword_config, num_keywords=num_keywords, budget=budget,
            render_mode=render_mode, loss_threshold=loss_threshold, max_days=max_days,
            updater_params=updater_params, updater_mask=updater_mask,
            obs_dtype=torch.float64, autoreset=False, n_lanes=1, detail_cap=kwargs.pop("detail_cap", 4096),
            **kwargs)
        self.keyword_config = keyword_config
        self.num_keywords = num_keywords
        self.budget = budget
        self.action_space = get_action_space(num_keywords)
        self.observation_space = get_observation_space(num_keywords, budget)
        self.render_mode = render_mode
        self.updater_params = updater_params
        self.updater_mask = updater_mask
        self._current_text = "New start\n"
        self._have_keywords = False
        self.current_day = 0
        self.cumulative_profit = 0.0

    # attributes the reference's callers read
    @property
    def max_days(self):
        return self._vec.max_days

    @max_days.setter
    def max_days(self, value):  # a plain attribute in the reference (env:80-82): callers assign it
        self._vec.max_days = int(value)

    @property
    def loss_threshold(self):
        return self._vec.loss_threshold

    @loss_threshold.setter
    def loss_threshold(self, value):
        self._vec.loss_threshold = float(value)

    @property
    def np_random(self):
        return self._vec.np_random

    @property
    def keywords(self):
        return self._vec.keywords

    @property
    def keyword_params(self) -> List[list]:
        """[[(vol_mean, vol_std), p1, p2|1/scale, bctr, sctr, mean_rev, std_rev], ...] with the
        drifted values (gymnasium_kw_utils.py:20-28)."""
        p = {k: v.reshape(-1) for k, v in self._vec.keyword_params().items()}
        implicit = self._vec.kind == 0
        return [[(p["vol_mean"][k], p["vol_std"][k]), p["p1"][k],
                 (1.0 / p["p2"][k]) if implicit else p["p2"][k], p["ctr"][k], p["cvr"][k],
                 p["rev_mean"][k], p["rev_std"][k]] for k in range(self.num_keywords)]

    def set_updater_mask(self, new_updater_mask: List[bool]) -> None:
        self._vec.set_updater_mask(new_updater_mask)
        self.updater_mask = new_updater_mask

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        self._vec.reset(seed=seed, options=options)
        # the reference reprs the CURRENT (possibly drifted) parameters (env:343-345)
        info = {"keyword_params": self._describe_params()}
        self._have_keywords = True
        self.current_day, self.cumulative_profit = 0, 0.0
        self._current_text = "Reset environment\n\nNew start\n"
        K = self.num_keywords
        # env:340-342: observation_space.sample() * 0 -> dtypes of the space (int64 / float32)
        obs = dict(impressions=np.zeros(K, np.int64), buyside_clicks=np.zeros(K, np.int64),
                   cost=np.zeros(K, np.float32), sellside_conversions=np.zeros(K, np.int64),
                   revenue=np.zeros(K, np.float32), cumulative_profit=np.zeros(1, np.float32),
                   days_passed=np.zeros(1, np.float32))
        return obs, info

    def step(self, action: Dict, *, tape=None):
        """``tape`` (a DeviceTape for this one env; parity mode): the step consumes pre-drawn
        volumes / competitor bids / uniforms / revenues instead of Philox draws."""
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        budget_in = action.get("budget", self.budget)
        # an ndarray budget is decremented in place by every lane AND by the campaign loop
        # (bidding_simulation.py:102 + :225): reproduce that double charge
        self._vec.budget_alias = isinstance(budget_in, np.ndarray) and budget_in.ndim >= 1
        raw = np.asarray(action["keyword_bids"])
        # numpy >= 2 keeps a float32 bid float32 through env:215 (the Box action space's dtype): such
        # a bid wins ties when its float32 value lies above the cent value
        f32 = raw.dtype == np.float32
        self._vec.f32_ties = bool(f32)
        bids_in = raw.astype(np.float32 if f32 else np.float64).reshape(1, -1)
        act = {"keyword_bids": bids_in,
               "budget": np.asarray(budget_in, dtype=np.float64).reshape(-1)[:1]}
        # the exact serial walk also records the per-click lists of info["bidding_outcomes"]
        if tape is not None:
            obs, reward, term, trunc, _ = self._vec.step_replay(act, tape, force_serial=True)
        else:
            obs, reward, term, trunc, _ = self._vec.step(act, force_serial=True)
        torch.cuda.current_stream(self._vec.device).synchronize()
        o = {k: v[0].cpu().numpy() for k, v in obs.items()}
        observations = dict(
            impressions=o["impressions"].astype(np.int64), buyside_clicks=o["buyside_clicks"].astype(np.int64),
            cost=o["cost"].astype(np.float64), sellside_conversions=o["sellside_conversions"].astype(np.int64),
            revenue=o["revenue"].astype(np.float64),
            cumulative_profit=o["cumulative_profit"].astype(np.float64).reshape(1),
            days_passed=o["days_passed"].astype(np.int64).reshape(1))
        self.current_day = int(observations["days_passed"][0])
        left = float(self._vec._out["remaining_budget"][0])
        self.budget = (np.array([left]) if self._vec.budget_alias
                       else np.round(np.asarray(budget_in, dtype=float), 2))
        bids = [float(np.round(np.maximum(b, 0.01), 2)) for b in bids_in[0]]  # in the bids' own dtype
        outcomes = self._vec.bidding_outcomes(0)
        for o_k, b in zip(outcomes, bids):
            o_k["bid"] = b
        # The device sums money as exact integer cents; the reference adds float64 dollars one click
        # at a time.  With every click's cost and revenue at hand the single-env adapter restates the
        # reference's own sums in its order (env:222-241: rust.sum_list over the concatenated lists,
        # profits lane by lane), so these floats are bit-identical to the reference's, not just
        # within the 1e-6 tolerance.
        observations["cost"] = np.array([_sum_list(k["costs"]) for k in outcomes])
        observations["revenue"] = np.array([_sum_list(k["revenues"]) for k in outcomes])
        profits = _sum_list(k["profit"] for k in outcomes)
        self.cumulative_profit += profits
        observations["cumulative_profit"] = np.array([self.cumulative_profit])
        info = {"bids": bids, "bidding_outcomes": repr_outcomes(outcomes),
                "keyword_params": self._describe_params()}
        terminated = bool(term[0])
        truncated = bool(self.cumulative_profit < -self.loss_threshold)
        # keep the device's episode state on the same float as the host's
        self._vec._state["cum_profit"][0] = self.cumulative_profit
        if self.render_mode == "ansi":
            self._current_text = (
                f"Time step: {self.current_day}/{self.max_days},   "
                + f"Average profit per kw in step: {profits / self.num_keywords:.2f},   "
                + f"Budget: {self.budget}   " + f"Total profit in step: {profits:.2f},   "
                + f"Cumulative profit: {self.cumulative_profit:.2f}\n")
        if truncated:
            self._current_text += (
                "Bidding simulation truncated early, we spent too much.\n"
                + f"Our allowed spend was ({self.loss_threshold:.2f}),\n"
                + f"but our cumulative loss was ({self.cumulative_profit:.2f})")
        return observations, profits, terminated, truncated, info

    def _describe_params(self) -> str:
        names = ["volume", "imp_intercept", "imp_slope", "bctr", "sctr", "mean revenue", "std revenue"]
        return "\n".join(f"kw{n} params:\n " + ",   ".join(f"{a}: {v}" for a, v in zip(names, p))
                         for n, p in enumerate(self.keyword_params))

    def render(self) -> Optional[str]:
        if self.render_mode == "ansi":
            return self._current_text
        return None

    def close(self) -> None:
        pass


def _rust_display(v: float) -> str:
    """Rust's ``{}`` for f64 (src/lib.rs:269): shortest round-trip digits, ALWAYS positional (no
    exponent at any magnitude), no trailing ``.0``."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    from decimal import Decimal
    s = format(Decimal(repr(v)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return "-0" if s in ("-", "-0") else s


def _rust_debug(v: float) -> str:
    """Rust's ``{:?}`` for f64 (the elements of ``{:?}`` on Vec<f64>, src/lib.rs:269): positional with
    at least one fractional digit for 0 and 1e-4 <= |v| < 1e16, otherwise shortest-digit scientific
    notation written like ``1.5e-7`` -- the same switch-over points as Python's repr, whose
    exponent is zero-padded and signed (``1.5e-07``, ``1e+16``)."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    s = repr(v)
    if "e" in s:
        mant, exp = s.split("e")
        if mant.endswith(".0"):
            mant = mant[:-2]
        return f"{mant}e{int(exp)}"
    return s


def _rust_debug_list(vs) -> str:
    return "[" + ", ".join(_rust_debug(v) for v in vs) + "]"


def repr_outcomes(outcomes: List[dict]) -> str:
    """``rust.repr_outcomes_py`` (src/lib.rs:250-275), host-side string formatting only."""
    parts = []
    for o in outcomes:
        parts.append(
            "{" + f"'bid': {_rust_display(o['bid'])}, 'impressions': {o['impressions']}, "
            f"'impression_share': {_rust_display(o['impression_share'])}, "
            f"'buyside_clicks': {o['buyside_clicks']}, 'costs': {_rust_debug_list(o['costs'])}, "
            f"'sellside_conversions': {o['sellside_conversions']}, "
            f"'revenues': {_rust_debug_list(o['revenues'])}, "
            f"'revenues_per_cost': {_rust_debug_list(o['revenues_per_cost'])}, "
            f"'profit': {_rust_display(o['profit'])}" + "}")
    return "[" + ", ".join(parts) + "]"


def bidding_sim_creator(env_config: Dict) -> BiddingSimulation:
    """gymnasium_kw_env.py:361-363."""
    return BiddingSimulation(**env_config)
