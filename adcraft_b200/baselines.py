"""Vectorised baseline bidder: ``NaiveZeroMarginStrategy`` of
``adcraft/baselines/interpolated_expectations.py:442-515`` for [E, K] (env, keyword) pairs at once.

Per keyword the reference keeps a cache {ave_rpc, num_rpc_obs, ave_sctr, num_sctr_obs} updated by
``update_cached_rpc_and_sctr`` (:105-152) from the last observation, and a ramp-up bid
``max_bids``.  The same arithmetic is done here on tensors (any device); the policy is not the hot
path, so plain torch ops are fine.  Reference quirks kept: the conversion-rate average is weighted
by the click count but its observation counter advances by one per step with clicks
(:100-101,146-151); with no conversion-rate observations the ramp-up test ``u <= 1/sqrt(0)`` is
always true (:502).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class VectorNaiveZeroMarginStrategy:
    def __init__(self, num_envs: int, num_keywords: int, default_expected_revenue_per_conversion: float = 3.0,
                 device="cpu", seed: Optional[int] = None, dtype=torch.float64):
        E, K = num_envs, num_keywords
        z = lambda v=0.0: torch.full((E, K), v, dtype=dtype, device=device)
        # get_empty_cache (:287-296)
        self.ave_rpc, self.num_rpc_obs = z(), z()
        self.ave_sctr, self.num_sctr_obs = z(0.4), z()
        self.max_bids = z(0.01)
        self.default_rpc = float(default_expected_revenue_per_conversion)
        self.prev_bids = None
        self.gen = torch.Generator(device=device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    # ------------------------------------------------------------------ cache update
    def update_all_caches(self, prev_action: Dict[str, torch.Tensor], prev_observation: Dict[str, torch.Tensor]) -> None:
        dt = self.ave_rpc.dtype
        self.prev_bids = prev_action["keyword_bids"].to(dt)
        clicks = prev_observation["buyside_clicks"].to(dt)
        conv = prev_observation["sellside_conversions"].to(dt)
        revenue = prev_observation["revenue"].to(dt)
        has_clicks = clicks > 0
        has_conv = has_clicks & (conv > 0)
        # ---- revenue per paid conversion (:126-140, process_rpc_and_update_cache :68-86)
        ave_rpc_new = revenue / conv  # used only where has_conv
        n_new = has_conv.to(dt)       # both "new_obs" and "num_rpc_obs" of a one-observation update
        tot = self.num_rpc_obs + n_new
        rpc = (ave_rpc_new * n_new + self.ave_rpc * self.num_rpc_obs) / torch.clamp(tot, min=1.0)
        self.ave_rpc = torch.where(has_conv, rpc, self.ave_rpc)
        self.num_rpc_obs = torch.where(has_conv, tot, self.num_rpc_obs)
        # ---- paid conversions per click (:120-124, process_sctr_and_update_cache :89-102)
        ave_sctr_new = torch.clamp(conv / clicks, min=0.0)
        n_clicks = torch.where(has_clicks, clicks, torch.zeros_like(clicks))
        all_obs = n_clicks + self.num_sctr_obs
        all_convs = ave_sctr_new * n_clicks + self.ave_sctr * self.num_sctr_obs
        sctr = all_convs / torch.clamp(all_obs, min=1.0)
        self.ave_sctr = torch.where(has_clicks, sctr, self.ave_sctr)
        self.num_sctr_obs = torch.where(has_clicks, self.num_sctr_obs + 1.0, self.num_sctr_obs)

    # ------------------------------------------------------------------ action
    def expected_rev_per_buyside_click(self) -> torch.Tensor:
        """get_expected_rev_per_buyside_click (:178-199), empirical fallbacks 0.3 / 0.7 (:168-175)."""
        none = (self.num_rpc_obs < 1) & (self.num_sctr_obs < 1)
        no_rpc = self.num_rpc_obs < 1
        return torch.where(none, torch.full_like(self.ave_rpc, 0.3),
                           torch.where(no_rpc, 0.7 * self.ave_sctr, self.ave_rpc * self.ave_sctr))

    def sample_action(self, uniforms: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Bids [E, K] and budget [E] (:494-515).  ``uniforms`` ([E, K], for tests) replaces the draws
        the reference takes for the keywords without a revenue observation."""
        assert self.prev_bids is not None, "call update_all_caches first"
        u = uniforms if uniforms is not None else torch.rand(
            self.ave_rpc.shape, dtype=self.ave_rpc.dtype, device=self.ave_rpc.device, generator=self.gen)
        no_rpc = self.num_rpc_obs < 1
        ramp = no_rpc & (u <= 1.0 / torch.sqrt(self.num_sctr_obs))  # 1/sqrt(0) = inf: always ramps
        stepped = torch.clamp(self.max_bids + 0.03, min=0.01, max=3.0)
        self.max_bids = torch.where(ramp, stepped, self.max_bids)
        bids = torch.where(ramp, stepped,
                           torch.where(no_rpc, self.ave_sctr * self.default_rpc, self.expected_rev_per_buyside_click()))
        weight = torch.where(ramp, 1.0, torch.where(no_rpc, 2.0, 3.0)).to(bids.dtype)
        return {"budget": 100.0 * weight.sum(dim=1), "keyword_bids": bids}
