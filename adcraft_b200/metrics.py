"""AKNCP / NCP metrics and their reduction across ranks.

``compute_AKNCP`` / ``compute_NCP`` follow ``adcraft/experiment_utils/experiment_metrics.py:64-83``
(inputs ``[T, K]``); the batched forms take ``[T, E, K]`` or running sums.  The ideal-profit
estimator follows ``experiment_metrics.py:20-61`` (2048 sampled competitor bids -> sort ->
searchsorted -> running mean; expected profit = vol_mean * impression_rate * bctr *
(sctr * mean_rev - cpc), clipped at 0, maximised over the bid grid).

The only collective on the path: ``reduce_metrics`` all-reduces a small float64 vector
(NCCL for CUDA tensors, gloo for CPU tensors).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch


def compute_AKNCP(kw_profits, ideal_profits) -> float:
    """Median over keywords of (time-mean profit / time-mean ideal profit), ideal <= 0 -> 1."""
    kw_profits, ideal_profits = np.asarray(kw_profits, np.float64), np.asarray(ideal_profits, np.float64)
    den = ideal_profits.copy()
    den[den <= 0] = 1.0
    den = den.mean(axis=0)
    return float(np.median(kw_profits.mean(axis=0) / den))


def compute_NCP(kw_profits, ideal_profits) -> float:
    den = float(np.asarray(ideal_profits).sum())
    if den <= 0.0:
        den = 1.0
    return float(np.asarray(kw_profits).sum() / den)


def implicit_bid_profile(loc: torch.Tensor, scale: torch.Tensor, bid_grid: torch.Tensor,
                         n_samples: int = 2048, generator: Optional[torch.Generator] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched ``get_implicit_kw_bid_cpc_impressions`` for [..., K] keywords on any device.

    Returns (impression_rate, expected_cpc), each ``[..., K, len(bid_grid)]``."""
    shape = loc.shape + (n_samples,)
    u = torch.rand(shape, dtype=torch.float64, device=loc.device, generator=generator)
    lap = torch.where(u >= 0.5, -torch.log(2.0 - 2.0 * u), torch.log(2.0 * u))  # numpy's laplace form
    bids = torch.round(torch.clamp((loc[..., None] + scale[..., None] * lap).abs(), min=0.0) * 100.0) / 100.0
    second, _ = torch.sort(bids, dim=-1)
    grid = bid_grid.to(torch.float64).expand(loc.shape + (bid_grid.numel(),)).contiguous()
    idx = torch.searchsorted(second, grid, right=True)
    rate = idx.to(torch.float64) / n_samples
    idx = torch.clamp(idx, max=n_samples - 1)
    mean_prices = torch.cumsum(second, dim=-1) / torch.arange(1, n_samples + 1, device=loc.device, dtype=torch.float64)
    return rate, torch.gather(mean_prices, -1, idx)


def max_expected_bid_profits(vol_mean, bctr, sctr, mean_rev, cpc, rate):
    """``get_max_expected_bid_profits`` (experiment_metrics.py:40-61), batched over [..., K]."""
    exp = torch.clamp(vol_mean[..., None] * rate * bctr[..., None] * (sctr[..., None] * mean_rev[..., None] - cpc),
                      min=0.0)
    best, arg = exp.max(dim=-1)
    return torch.clamp(best, min=0.0), (exp > 0).sum(-1).to(torch.float64) / exp.shape[-1], arg


class MetricAccumulator:
    """Running sums over an episode of per-keyword profit and ideal profit, on the env's device."""

    def __init__(self, num_envs: int, num_keywords: int, device):
        z = lambda: torch.zeros(num_envs, num_keywords, dtype=torch.float64, device=device)
        self.kw_profit, self.ideal, self.ideal_den = z(), z(), z()
        self.steps = 0
        self.reward_sum = torch.zeros((), dtype=torch.float64, device=device)
        self.episodes = torch.zeros((), dtype=torch.float64, device=device)

    def update(self, obs: Dict[str, torch.Tensor], reward: torch.Tensor, ideal: Optional[torch.Tensor] = None,
               done: Optional[torch.Tensor] = None) -> None:
        self.kw_profit += (obs["revenue"] - obs["cost"]).to(torch.float64)
        if ideal is not None:
            ideal = ideal.to(torch.float64).expand_as(self.ideal)
            self.ideal += ideal
            self.ideal_den += torch.where(ideal <= 0, torch.ones_like(ideal), ideal)
        self.reward_sum += reward.sum()
        if done is not None:
            self.episodes += done.sum()
        self.steps += 1

    def per_env(self) -> Dict[str, torch.Tensor]:
        """AKNCP / NCP of every env over the accumulated steps (median over its keywords, local)."""
        t = max(self.steps, 1)
        akncp = torch.quantile((self.kw_profit / t) / (self.ideal_den / t), 0.5, dim=1)  # == np.median
        den = self.ideal.sum(1)
        ncp = self.kw_profit.sum(1) / torch.where(den <= 0, torch.ones_like(den), den)
        return {"akncp": akncp, "ncp": ncp}

    def summary_vector(self) -> torch.Tensor:
        """[sum kw_profit, sum ideal, sum akncp, sum akncp^2, sum ncp, n_env, sum reward, episodes]."""
        pe = self.per_env()
        return torch.stack([self.kw_profit.sum(), self.ideal.sum(), pe["akncp"].sum(), (pe["akncp"] ** 2).sum(),
                            pe["ncp"].sum(), torch.tensor(float(self.kw_profit.shape[0]), dtype=torch.float64,
                                                          device=self.kw_profit.device),
                            self.reward_sum, self.episodes])


def reduce_metrics(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum-all-reduce of the metric vector over ranks (no-op without an initialised group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def summarize(vec: torch.Tensor) -> Dict[str, float]:
    v = vec.detach().cpu().numpy()
    n = max(v[5], 1.0)
    mean = v[2] / n
    var = max(v[3] / n - mean * mean, 0.0)
    return dict(ncp_pooled=float(v[0] / (v[1] if v[1] > 0 else 1.0)), akncp_mean=float(mean),
                akncp_sem=float(np.sqrt(var / n)), ncp_mean=float(v[4] / n), n_envs=float(v[5]),
                reward_sum=float(v[6]), episodes=float(v[7]))
