"""Flat array views of observations / actions (``adcraft/wrappers/flat_array.py:10-87``,
``gymnasium_kw_utils.py:383-390``).

Key order is the sorted one gymnasium uses for Dict spaces:
observation ``[buyside_clicks(K) | cost(K) | cumulative_profit(1) | days_passed(1) |
impressions(K) | revenue(K) | sellside_conversions(K)]`` (5K+2), action ``[budget(1) | keyword_bids(K)]``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .spaces import Box

OBS_KEYS_SORTED = ("buyside_clicks", "cost", "cumulative_profit", "days_passed", "impressions",
                   "revenue", "sellside_conversions")


def flatten_dict_array(obs: Dict[str, np.ndarray]) -> np.ndarray:
    """gymnasium_kw_utils.py:383-390."""
    return np.hstack([np.asarray(obs[k]).ravel() for k in sorted(obs.keys())])


def flat_observations(obs, env=None):
    """[E, 5K+2] flat observation rows.  With ``env`` built with ``flat_obs=True`` this is the tensor
    the step kernels wrote (zero-copy, ``env.flat_observation()``); otherwise the rows are gathered
    from the observation dict (device, float of cost's dtype)."""
    import torch
    if env is not None and getattr(env, "want_flat_obs", False):
        return env.flat_observation()
    dt = obs["cost"].dtype
    return torch.cat([obs[k].to(dt).reshape(obs[k].shape[0], -1) for k in OBS_KEYS_SORTED], dim=1)


def unflatten_actions(flat):
    """[E, K+1] -> {"budget": [E], "keyword_bids": [E, K]} (views, no copy)."""
    return {"budget": flat[:, 0], "keyword_bids": flat[:, 1:]}


def observation_slices(num_keywords: int) -> Dict[str, slice]:
    K, out, pos = num_keywords, {}, 0
    for k in OBS_KEYS_SORTED:
        n = 1 if k in ("cumulative_profit", "days_passed") else K
        out[k] = slice(pos, pos + n)
        pos += n
    return out


class FlatArrayWrapper:
    """Flattens observations and actions of the single-env adapter to Box spaces."""

    def __init__(self, env):
        self.env = env
        K = env.num_keywords
        lo = np.concatenate([np.zeros(2 * K), [-np.inf, 0.0], np.zeros(3 * K)])
        hi = np.full(5 * K + 2, np.inf)
        hi[K:2 * K] = float(np.asarray(env.budget).ravel()[0])
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(5 * K + 2,), dtype=np.float64)
        self.observation_space.low, self.observation_space.high = lo, hi
        self.action_space = Box(low=0.01, high=np.inf, shape=(K + 1,), dtype=np.float32)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def action(self, action: np.ndarray) -> Dict[str, np.ndarray]:
        a = np.asarray(action)
        return {"budget": a[:1], "keyword_bids": a[1:]}

    def observation(self, observation: Dict[str, np.ndarray]) -> np.ndarray:
        return flatten_dict_array(observation)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(self.action(action))
        return flatten_dict_array(obs), reward, terminated, truncated, info

    def reset(self, *args, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, info = self.env.reset(*args, seed=seed, options=options)
        return self.observation(obs), info
