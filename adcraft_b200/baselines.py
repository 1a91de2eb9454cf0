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


class VectorNaiveInterpolationStrategy:
    """``NaiveInterpolationStrategy`` (``interpolated_expectations.py:298-439``) for [E, K] pairs.

    The reference keeps, per keyword, dicts ``{bid -> [running mean, count]}`` of cost-per-click and
    clicks plus the rpc / sctr cache of the zero-margin agent; to act it interpolates both dicts
    over the 300 allowed bids (values first smoothed over the *observed* bids with a Bartlett
    window, :208-216), turns margin above a threshold into a distribution truncated one
    ``bid_step`` above the highest bid tried so far, and samples a bid from it (:367-439).

    Here the dicts are dense ``[E, K, 300]`` tables over the bid grid (a bid is its grid index
    ``round(100 bid) - 1``; the agent only ever bids grid values) and every step is a handful of
    tensor ops.  Quirks kept: running means use the reference's update formula and dtypes
    (cost-per-click in f64, clicks in f32, rpc / sctr in f32), the cpc table gets a first entry
    only from a step with clicks (:50-54), the clicks table from any step (:29-31), the smoothing
    window is 1 tap up to four observed bids, ``[.5, .5]`` trailing at five and ``[.25, .5, .25]``
    from six (``np.bartlett(min(5, max(1, n-1)))`` under ``np.convolve(mode="same")``), the
    no-data prior is ``cpc = 0.9 bid, clicks = 1`` keyed on the *cpc* table being empty (:247,266-270).
    """

    GRID = 300

    def __init__(self, num_envs: int, num_keywords: int, profit_acquisition_threshold: float = -0.2,
                 device="cpu", seed: Optional[int] = None, bid_step: float = 0.03):
        E, K, G = num_envs, num_keywords, self.GRID
        self.E, self.K = E, K
        self.device = torch.device(device)
        f32, f64 = torch.float32, torch.float64
        z = lambda shape, v=0.0, dt=f64: torch.full(shape, v, dtype=dt, device=device)
        self.ave_rpc, self.num_rpc_obs = z((E, K), 0.0, f32), z((E, K))
        self.ave_sctr, self.num_sctr_obs = z((E, K), 0.4, f32), z((E, K))
        self.ave_cpc, self.n_cpc = z((E, K, G)), z((E, K, G))
        self.ave_clicks, self.n_clicks = z((E, K, G), 0.0, f32), z((E, K, G))
        self.max_observed_bid = z((E, K), 0.03)  # observed_bids.append(0.03) (:386)
        self.profit_acquisition_threshold = float(profit_acquisition_threshold)
        self.bid_step = float(bid_step)
        import numpy as np
        self.allowed_bids = torch.from_numpy(np.linspace(0.01, 3.00, G)).to(device)          # x of np.interp
        self.grid_bids = torch.from_numpy(np.arange(0.01, 3.01, 0.01)[:G].copy()).to(device)  # xp of np.interp
        self.profit_beliefs = self.cost_beliefs = None
        self.gen = torch.Generator(device=device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    # ------------------------------------------------------------------ cache update (:219-241, :105-152)
    def update_all_caches(self, prev_action: Dict[str, torch.Tensor], prev_observation: Dict[str, torch.Tensor]) -> None:
        f32, f64 = torch.float32, torch.float64
        bids = prev_action["keyword_bids"].to(self.device)
        # observations pass through a float32 torch.Tensor in the reference (:352-354)
        clicks = prev_observation["buyside_clicks"].to(self.device).to(f32)
        conv = prev_observation["sellside_conversions"].to(self.device).to(f32)
        revenue = prev_observation["revenue"].to(self.device).to(f32)
        cost = prev_observation["cost"].to(self.device).to(f32)
        has_clicks = clicks > 0
        has_conv = has_clicks & (conv > 0)
        # rpc (:126-140, :68-86): one observation of revenue / conversions
        n_new = has_conv.to(f64)
        tot = self.num_rpc_obs + n_new
        rpc = ((revenue / conv) * n_new.to(f32) + self.ave_rpc * self.num_rpc_obs.to(f32)) / torch.clamp(tot, min=1.0).to(f32)
        self.ave_rpc = torch.where(has_conv, rpc, self.ave_rpc)
        self.num_rpc_obs = torch.where(has_conv, tot, self.num_rpc_obs)
        # sctr (:120-124, :89-102): weighted by this step's clicks, counter + 1
        n_clk = torch.where(has_clicks, clicks, torch.zeros_like(clicks))
        all_obs = n_clk.to(f64) + self.num_sctr_obs
        all_convs = torch.clamp(conv / clicks, min=0.0) * n_clk + self.ave_sctr * self.num_sctr_obs.to(f32)
        sctr = all_convs / torch.clamp(all_obs, min=1.0).to(f32)
        self.ave_sctr = torch.where(has_clicks, sctr, self.ave_sctr)
        self.num_sctr_obs = torch.where(has_clicks, self.num_sctr_obs + 1.0, self.num_sctr_obs)
        # per-bid running means (:22-65)
        g = torch.clamp(torch.round(bids.to(f64) * 100.0).to(torch.int64) - 1, 0, self.GRID - 1).unsqueeze(-1)
        cpc = cost.to(f64) / clicks.to(f64)                      # compute_cpc: python floats (:15-19)
        a, n = self.ave_cpc.gather(-1, g).squeeze(-1), self.n_cpc.gather(-1, g).squeeze(-1)
        new_a = torch.where(n > 0, (cpc + a * n) / (1.0 + n), cpc)
        self.ave_cpc.scatter_(-1, g, torch.where(has_clicks, new_a, a).unsqueeze(-1))
        self.n_cpc.scatter_(-1, g, torch.where(has_clicks, n + 1.0, n).unsqueeze(-1))
        clk0 = torch.where(has_clicks, clicks, torch.zeros_like(clicks))  # nan -> 0 (:228)
        a, n = self.ave_clicks.gather(-1, g).squeeze(-1), self.n_clicks.gather(-1, g).squeeze(-1)
        new_a = torch.where(n > 0, (clk0 + a * n.to(f32)) / (1.0 + n).to(f32), clk0)
        self.ave_clicks.scatter_(-1, g, new_a.unsqueeze(-1))
        self.n_clicks.scatter_(-1, g, (n + 1.0).unsqueeze(-1))
        self.max_observed_bid = torch.maximum(self.max_observed_bid, torch.round(bids.to(f64) * 100.0) / 100.0)

    # ------------------------------------------------------------------ beliefs (:155-285)
    def expected_rev_per_buyside_click(self) -> torch.Tensor:
        no_rpc, no_sctr = self.num_rpc_obs < 1, self.num_sctr_obs < 1
        learned = self.ave_rpc.to(torch.float64) * self.ave_sctr.to(torch.float64)
        return torch.where(no_rpc & no_sctr, torch.full_like(learned, 0.3),
                           torch.where(no_rpc, 0.7 * self.ave_sctr.to(torch.float64), learned))

    def _interp(self, ave: torch.Tensor, cnt: torch.Tensor, left, right_is_max: bool):
        """np.interp(allowed_bids, observed bids, smoothed(observed means), left, right) per pair."""
        f64 = torch.float64
        E, K, G = cnt.shape
        obs = cnt > 0
        n = obs.sum(-1)                                                     # observed bids per pair
        rank = torch.cumsum(obs, -1) - 1                                    # position among the observed
        v = torch.zeros(E, K, G + 2, dtype=f64, device=ave.device)          # compacted values, zero padded
        idx = torch.where(obs, rank + 1, torch.full_like(rank, G + 1))      # unobserved -> dump slot
        v.scatter_(-1, idx, ave.to(f64))
        v[..., G + 1] = 0.0
        xp = torch.zeros(E, K, G + 2, dtype=f64, device=ave.device)
        xp.scatter_(-1, idx, self.grid_bids.expand(E, K, G).contiguous())
        vm1, v0, vp1 = v[..., :G], v[..., 1:G + 1], torch.cat([v[..., 2:G + 1], torch.zeros_like(v[..., :1])], -1)
        nn = n.unsqueeze(-1)
        sm = torch.where(nn >= 6, (0.25 * vm1 + 0.5 * v0) + 0.25 * vp1, torch.where(nn == 5, 0.5 * vm1 + 0.5 * v0, v0))
        pos = torch.arange(G, device=ave.device).expand(E, K, G)
        sm = torch.where(pos < nn, sm, torch.zeros_like(sm))                # smoothed[j], j < n
        xs = xp[..., 1:G + 1]                                               # xp[j], j < n
        x = self.allowed_bids.expand(E, K, G)
        # j = index of the last observed bid <= x  (np.interp's binary search)
        big = torch.where(pos < nn, xs, torch.full_like(xs, float("inf")))
        j = torch.searchsorted(big.contiguous(), x.contiguous(), right=True) - 1
        jc = torch.clamp(j, 0, G - 2)
        f0, f1 = sm.gather(-1, jc), sm.gather(-1, jc + 1)
        x0, x1 = xs.gather(-1, jc), xs.gather(-1, jc + 1)
        lin = (f1 - f0) / (x1 - x0) * (x - x0) + f0
        last = torch.clamp(nn - 1, min=0)
        f_last, x_last = sm.gather(-1, last), xs.gather(-1, last)
        if right_is_max:  # right=np.max(ave_cpcs): of the RAW means (:256)
            raw = torch.where(obs, ave.to(f64), torch.full((), float("-inf"), dtype=f64, device=ave.device))
            right = raw.max(-1, keepdim=True).values
        else:
            right = f_last
        left_t = left if torch.is_tensor(left) else torch.full_like(lin, left)
        out = torch.where(j < 0, left_t,
                          torch.where(x > x_last, right.expand_as(lin),
                                      torch.where(j >= last, f_last.expand_as(lin),
                                                  torch.where(x == x0, f0, lin))))
        return out, n

    def expected_margins_and_costs(self):
        raw_first = None
        obs_c = self.n_clicks > 0
        first = torch.argmax(obs_c.to(torch.int8), -1, keepdim=True)
        raw_first = self.ave_clicks.gather(-1, first).to(torch.float64)      # left=ave_clicks[0] (raw, :262)
        last = (self.GRID - 1) - torch.argmax(obs_c.flip(-1).to(torch.int8), -1, keepdim=True)
        raw_last = self.ave_clicks.gather(-1, last).to(torch.float64)        # right=ave_clicks[-1] (raw, :263)
        cpc, n_cpc = self._interp(self.ave_cpc, self.n_cpc, 0.01, True)
        clk, _ = self._interp(self.ave_clicks, self.n_clicks, raw_first.expand(-1, -1, self.GRID), False)
        x = self.allowed_bids.expand_as(clk)
        x_last = self.grid_bids[last.squeeze(-1)].unsqueeze(-1)
        clk = torch.where(x > x_last, raw_last.expand_as(clk), clk)
        have = (n_cpc > 0).unsqueeze(-1)                                     # np.any(unique_bids_cpc) (:247)
        cpc = torch.where(have, cpc, 0.9 * x)
        clk = torch.where(have, clk, torch.ones_like(clk))
        rev = self.expected_rev_per_buyside_click().unsqueeze(-1)
        return (-cpc + rev) * (0.01 + clk), cpc * (0.01 + clk)

    def acquisition(self, margins: torch.Tensor):
        """(:367-398) probabilities over the grid and the pairs that have any mass."""
        thr = -(1.0 / (1.0 + self.num_rpc_obs + self.num_sctr_obs / 5.0)) * abs(self.profit_acquisition_threshold)
        acq = torch.clamp(margins, min=thr.unsqueeze(-1)) - thr.unsqueeze(-1)
        end = torch.clamp((100.0 * (self.max_observed_bid + self.bid_step) - 1.0).to(torch.int64), max=self.GRID)
        pos = torch.arange(self.GRID, device=margins.device).expand_as(acq)
        acq = torch.where(pos < end.unsqueeze(-1), acq, torch.zeros_like(acq))
        mass = acq.sum(-1, keepdim=True)
        return acq / torch.where(mass > 0, mass, torch.ones_like(mass)), (mass > 0).squeeze(-1)

    def sample_action(self, uniforms: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """(:405-439).  ``uniforms`` [E, K]: the draw ``Generator.choice(p=...)`` would consume for each
        pair that has mass (inverse-cdf sampling, ``cdf.searchsorted(u, side="right")``)."""
        margins, costs = self.expected_margins_and_costs()
        p, has_mass = self.acquisition(margins)
        if uniforms is None:
            uniforms = torch.rand(self.E, self.K, dtype=torch.float64, device=self.device, generator=self.gen)
        cdf = torch.cumsum(p, -1)
        cdf = cdf / torch.where(has_mass, cdf[..., -1], torch.ones_like(cdf[..., -1])).unsqueeze(-1)
        idx = torch.searchsorted(cdf.contiguous(), uniforms.to(self.device).to(torch.float64).unsqueeze(-1), right=True)
        idx = torch.clamp(idx, max=self.GRID - 1)
        bid = self.allowed_bids[idx.squeeze(-1)]
        bids = torch.where(has_mass, bid, torch.full_like(bid, 0.01))
        c_at = costs.gather(-1, idx).squeeze(-1)
        m_at = margins.gather(-1, idx).squeeze(-1)
        zero = torch.zeros_like(bid)
        exp_cost = torch.where(has_mass, torch.where(self.num_sctr_obs > 0, c_at, bid), zero).sum(-1)
        exp_profit = torch.where(has_mass & (self.num_rpc_obs > 0), m_at, zero).sum(-1)
        self.profit_beliefs, self.cost_beliefs = exp_profit, exp_cost
        base = torch.clamp(torch.clamp(exp_cost, max=10000.0), min=1000.0)
        budget = torch.where(exp_profit > 0, 1.5 * base,
                             torch.where(exp_profit > self.K * self.profit_acquisition_threshold, base,
                                         torch.full_like(base, 1000.0)))
        return {"budget": budget, "keyword_bids": bids}
