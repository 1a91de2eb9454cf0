"""AKNCP / NCP metrics and their reduction across ranks.

``compute_AKNCP`` / ``compute_NCP`` follow ``adcraft/experiment_utils/experiment_metrics.py:64-83``
(inputs ``[T, K]``); the batched forms take ``[T, E, K]`` or running sums.  The ideal-profit
estimator (``experiment_metrics.py:20-61``: 2048 sampled competitor bids -> sort -> searchsorted ->
running mean; expected profit = vol_mean * impression_rate * bctr * (sctr * mean_rev - cpc),
clipped at 0, maximised over the bid grid) is a CUDA kernel behind ``adc_ideal_profit``
(csrc/adc_metrics.cu); ``ideal_profit`` below is its host-side call.

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


DEFAULT_BID_GRID = np.arange(0.01, 3.00, 0.01)  # the notebooks' allowed_bids (run_heatmap_experiments cell 3)


def ideal_profit(env, bid_grid: Optional[np.ndarray] = None, n_samples: int = 2048, *,
                 samples_cents: Optional[torch.Tensor] = None, step: Optional[int] = None,
                 profile: bool = False) -> Dict[str, torch.Tensor]:
    """``get_implicit_kw_bid_cpc_impressions`` + ``get_max_expected_bid_profits``
    (experiment_metrics.py:20-61) for every (env, keyword) of a VectorBiddingSimulation, on its
    device, through the C ABI (``adc_ideal_profit``: counting sort of the sampled competitor bids
    in shared memory, one warp per unit).  Reads the env's CURRENT (drifted) parameters.

    A keyword set shared by all envs gives ``[1, K]`` results (broadcast them), per-env sets
    ``[E, K]``.  ``samples_cents`` ([rows, K, n_samples] int32 on the device) replaces the Philox
    draws with given competitor bids (parity with the reference on its own samples).
    Returns ``ideal`` (max expected profit), ``positive_frac``, ``best_bid_index`` and, with
    ``profile``, ``impression_rate`` / ``expected_cpc`` of shape ``[rows, K, len(bid_grid)]``."""
    import ctypes as C
    from . import _capi
    grid = np.ascontiguousarray(DEFAULT_BID_GRID if bid_grid is None else bid_grid, dtype=np.float64)
    dev, K = env.device, env.num_keywords
    rows = env.num_envs if env._kw_stride else 1
    a = _capi.IdealArgs()
    a.E, a.env_base, a.seed = rows, env.env_base, env.seed & 0xFFFFFFFFFFFFFFFF
    a.step = (env._step_count if step is None else step) & 0xFFFFFFFF
    a.device = dev.index
    kw = a.kw
    kw.kind, kw.K, kw.env_stride = env.kind, K, env._kw_stride
    for n in ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std"):
        setattr(kw, n, env._kw_dev[n].data_ptr())
    a.n_samples, a.n_grid, a.bid_grid_host = int(n_samples), len(grid), grid.ctypes.data
    if samples_cents is not None:
        assert samples_cents.dtype == torch.int32 and samples_cents.is_contiguous() and samples_cents.device == dev
        assert tuple(samples_cents.shape) == (rows, K, n_samples)
        a.samples_cents = samples_cents.data_ptr()
    f64 = torch.float64
    out = dict(ideal=torch.empty(rows, K, dtype=f64, device=dev), positive_frac=torch.empty(rows, K, dtype=f64, device=dev),
               best_bid_index=torch.empty(rows, K, dtype=torch.int32, device=dev))
    a.ideal_profit, a.positive_frac = out["ideal"].data_ptr(), out["positive_frac"].data_ptr()
    a.best_bid_index = out["best_bid_index"].data_ptr()
    if profile:
        out["impression_rate"] = torch.empty(rows, K, len(grid), dtype=f64, device=dev)
        out["expected_cpc"] = torch.empty(rows, K, len(grid), dtype=f64, device=dev)
        a.impression_rate, a.expected_cpc = out["impression_rate"].data_ptr(), out["expected_cpc"].data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    env._call(env._lib.adc_ideal_profit, C.byref(a), C.c_void_p(stream))
    return out


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


def episode_summary_vector(env, steps: int, ideal: torch.Tensor, reward_sum: Optional[torch.Tensor] = None,
                           episodes: Optional[torch.Tensor] = None, zero: bool = True,
                           row_chunk: int = 8192, use_kernel: bool = True) -> torch.Tensor:
    """The metric vector of ``MetricAccumulator.summary_vector`` from the env's own kernel-side
    accumulators (``episode_profit=True``: ``adc_step_out.episode_profit_cents`` holds the exact sum
    of every step's per-keyword profit): per-env AKNCP = median over keywords of
    (mean profit / mean ideal, ideal <= 0 -> 1) and NCP = sum profit / sum ideal
    (experiment_metrics.py:64-83) over ``steps`` steps with a per-step ideal profit ``ideal``
    ([1, K] or [E, K], stationary over the window).  ``zero``: start the next window."""
    from . import _capi
    acc = env._out["episode_profit_cents"]
    E, K = acc.shape
    f64 = torch.float64
    ideal = ideal.to(f64)
    out = torch.zeros(8, dtype=f64, device=acc.device)
    er, ec = env._out.get("episode_reward"), env._out.get("episode_count")
    if use_kernel and acc.is_cuda and K <= _capi.METRICS_MAX_K and ideal.dim() == 2 and ideal.shape[0] in (1, E):
        # one launch: a warp per env (adc_episode_metrics); the torch form below is its reference
        import ctypes as C
        idl = ideal.contiguous()
        a = _capi.MetricsArgs()
        a.E, a.K, a.steps, a.device = E, K, int(steps), acc.device.index
        a.episode_profit_cents, a.ideal = acc.data_ptr(), idl.data_ptr()
        a.ideal_env_stride = 0 if idl.shape[0] == 1 else K
        a.sums, a.akncp, a.ncp, a.zero = out.data_ptr(), None, None, int(zero)
        with torch.cuda.device(acc.device):
            _capi.check(_capi.load().adc_episode_metrics(C.byref(a), C.c_void_p(torch.cuda.current_stream(acc.device).cuda_stream)))
        out[6] = reward_sum if reward_sum is not None else (er.sum() if er is not None else 0.0)
        out[7] = episodes if episodes is not None else (ec.sum() if ec is not None else 0.0)
        if zero and er is not None:
            er.zero_()
            ec.zero_()
        return out
    for r0 in range(0, E, row_chunk):
        prof = acc[r0:r0 + row_chunk].to(f64) / 100.0
        idl = ideal if ideal.shape[0] == 1 else ideal[r0:r0 + row_chunk]
        den = torch.where(idl <= 0, torch.ones_like(idl), idl)
        ratio = (prof / steps) / den
        # np.median: mean of the two middle values for an even count
        lo = torch.kthvalue(ratio, (K + 1) // 2, dim=1).values
        hi = torch.kthvalue(ratio, K // 2 + 1, dim=1).values
        akncp = 0.5 * (lo + hi)
        isum = idl.sum(1) * steps
        ncp = prof.sum(1) / torch.where(isum <= 0, torch.ones_like(isum), isum)
        out[0] += prof.sum()
        out[1] += (idl.expand_as(prof) * steps).sum()
        out[2] += akncp.sum()
        out[3] += (akncp ** 2).sum()
        out[4] += ncp.sum()
        out[5] += prof.shape[0]
    out[6] = reward_sum if reward_sum is not None else (er.sum() if er is not None else 0.0)
    out[7] = episodes if episodes is not None else (ec.sum() if ec is not None else 0.0)
    if zero:
        acc.zero_()
        if er is not None:
            er.zero_()
            ec.zero_()
    return out


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
