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
from .vector_env import DEFAULT_UPDATER_PARAMS, VectorBiddingSimulation

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

    @property
    def loss_threshold(self):
        return self._vec.loss_threshold

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
        _, info = self._vec.reset(seed=seed, options=options)
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

    def step(self, action: Dict):
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        budget_in = action.get("budget", self.budget)
        # an ndarray budget is decremented in place by every lane AND by the campaign loop
        # (bidding_simulation.py:102 + :225): reproduce that double charge
        self._vec.budget_alias = isinstance(budget_in, np.ndarray) and budget_in.ndim >= 1
        bids_in = np.asarray(action["keyword_bids"], dtype=np.float64).reshape(1, -1)
        act = {"keyword_bids": bids_in,
               "budget": np.asarray(budget_in, dtype=np.float64).reshape(-1)[:1]}
        # the exact serial walk also records the per-click lists of info["bidding_outcomes"]
        obs, reward, term, trunc, _ = self._vec.step(act, force_serial=True)
        torch.cuda.current_stream(self._vec.device).synchronize()
        o = {k: v[0].cpu().numpy() for k, v in obs.items()}
        observations = dict(
            impressions=o["impressions"].astype(np.int64), buyside_clicks=o["buyside_clicks"].astype(np.int64),
            cost=o["cost"].astype(np.float64), sellside_conversions=o["sellside_conversions"].astype(np.int64),
            revenue=o["revenue"].astype(np.float64),
            cumulative_profit=o["cumulative_profit"].astype(np.float64).reshape(1),
            days_passed=o["days_passed"].astype(np.int64).reshape(1))
        profits = float(reward[0])
        self.cumulative_profit = float(observations["cumulative_profit"][0])
        self.current_day = int(observations["days_passed"][0])
        left = float(self._vec._out["remaining_budget"][0])
        self.budget = (np.array([left]) if self._vec.budget_alias
                       else np.round(np.asarray(budget_in, dtype=float), 2))
        bids = [float(np.round(np.maximum(b, 0.01), 2)) for b in bids_in[0]]
        outcomes = self._vec.bidding_outcomes(0)
        info = {"bids": bids, "bidding_outcomes": repr_outcomes(outcomes),
                "keyword_params": self._describe_params()}
        terminated, truncated = bool(term[0]), bool(trunc[0])
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
    """Rust's `{}` for f64 (src/lib.rs:269): shortest round-trip digits, no trailing `.0`."""
    v = float(v)
    if v == int(v) and abs(v) < 1e16:
        return str(int(v))
    return repr(v)


def _rust_debug_list(vs) -> str:
    """Rust's `{:?}` for Vec<f64>: always a decimal point."""
    return "[" + ", ".join(repr(float(v)) for v in vs) + "]"


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
