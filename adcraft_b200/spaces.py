"""Action / observation spaces of the env (``adcraft/gymnasium_kw_utils.py:31-64``).

Uses gymnasium's ``Box`` / ``Dict`` when gymnasium is importable; otherwise small duck-typed
stand-ins with the same ``shape / dtype / low / high / sample / contains`` surface, so the env
can be constructed in images without gymnasium (this one has none).
"""
from __future__ import annotations

from typing import Dict as _Dict

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym  # type: ignore
    if getattr(_gym, "__file__", None) is None:  # an in-memory stand-in (the test harness installs one)
        raise ImportError("gymnasium is a stub module")
    from gymnasium.spaces import Box, Dict  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:  # minimal gymnasium.spaces.Box
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=np.float64)
            self.high = np.full(self.shape, high, dtype=np.float64)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            finite_hi = np.isfinite(self.high)
            x = np.where(finite_hi, self._rng.uniform(lo, np.where(finite_hi, self.high, lo + 1.0)),
                         lo + self._rng.exponential(size=self.shape))
            return x.astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:  # minimal gymnasium.spaces.Dict (keys kept in sorted order like gymnasium)
        def __init__(self, spaces: _Dict[str, Box]):
            self.spaces = dict(sorted(spaces.items()))

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x) -> bool:
            return (isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys())
                    and all(self.spaces[k].contains(v) for k, v in x.items()))

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k}: {v}" for k, v in self.spaces.items()) + ")"


def get_action_space(num_keywords: int) -> Dict:
    """gymnasium_kw_utils.py:31-42."""
    return Dict({
        "keyword_bids": Box(low=0.01, high=float("Inf"), shape=(num_keywords,), dtype=np.float32),
        "budget": Box(low=0.01, high=float("Inf"), shape=(1,), dtype=np.float32),
    })


def get_observation_space(num_keywords: int, budget: float) -> Dict:
    """gymnasium_kw_utils.py:45-64."""
    nonneg_int = lambda: Box(low=0, high=float("Inf"), shape=(num_keywords,), dtype=int)
    return Dict({
        "impressions": nonneg_int(),
        "buyside_clicks": nonneg_int(),
        "cost": Box(low=0, high=budget, shape=(num_keywords,), dtype=np.float32),
        "sellside_conversions": nonneg_int(),
        "revenue": Box(low=0, high=float("Inf"), shape=(num_keywords,), dtype=np.float32),
        "cumulative_profit": Box(low=-float("Inf"), high=float("Inf"), shape=(1,), dtype=np.float32),
        "days_passed": Box(low=0, high=float("Inf"), shape=(1,), dtype=np.float32),
    })
