"""Multi-agent view (``adcraft/multi_agent/env.py:8-35``).

The reference builds its multi-agent env with RLlib's ``make_multi_agent``: ``num_agents``
*independent* ``FlatArrayWrapper(BiddingSimulation())`` copies stepped with ``{agent_id: action}``
dicts -- there is no shared auction.  Here the A agents of each of the
E "worlds" are A*E rows of one VectorBiddingSimulation (row = world * A + agent, each with its own
keyword set like the reference's independently constructed copies), so a step of all agents is
still one launch.  Flat observations / actions use the reference's sorted-key layout.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .vector_env import VectorBiddingSimulation
from .wrappers import flat_observations, unflatten_actions


class MultiAgentBiddingSimulation:
    def __init__(self, num_agents: int, num_worlds: int = 1, **env_kwargs):
        self.num_agents, self.num_worlds = int(num_agents), int(num_worlds)
        env_kwargs.setdefault("shared_keywords", False)  # independent copies draw their own keywords
        self.vec = VectorBiddingSimulation(self.num_agents * self.num_worlds, **env_kwargs)
        self._agent_ids = list(range(self.num_agents))

    def get_agent_ids(self):
        return set(self._agent_ids)

    def _split(self, t: torch.Tensor) -> Dict[int, torch.Tensor]:
        v = t.view(self.num_worlds, self.num_agents, *t.shape[1:])
        return {a: v[:, a] for a in self._agent_ids}

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, info = self.vec.reset(seed=seed, options=options)
        return self._split(flat_observations(obs)), {a: info for a in self._agent_ids}

    def step(self, action_dict: Dict[int, torch.Tensor]):
        """action_dict[agent] = flat actions [worlds, K+1] = [budget | keyword_bids] (all agents)."""
        flat = torch.stack([action_dict[a] for a in self._agent_ids], dim=1)
        flat = flat.reshape(self.num_agents * self.num_worlds, -1).contiguous()
        act = unflatten_actions(flat)
        obs, reward, term, trunc, info = self.vec.step(
            {"keyword_bids": act["keyword_bids"].contiguous(), "budget": act["budget"].contiguous()})
        o, r, te, tr = (self._split(x) for x in (flat_observations(obs), reward, term, trunc))
        te["__all__"] = term.view(self.num_worlds, self.num_agents).all(dim=1)
        tr["__all__"] = trunc.view(self.num_worlds, self.num_agents).any(dim=1)
        return o, r, te, tr, {a: info for a in self._agent_ids}


class SharedAuctionSimulation:
    """A bidders competing inside ONE auction per (world, keyword, auction) -- BASELINE config 4
    ("8 competing bidders per auction sharing keyword volume").  The reference
    has no such mode (its multi-agent env is independent copies, above); the semantics are the
    reference's own auction applied to the obvious ``other_bids`` matrix: bidder a faces the A-1
    rival bids plus the keyword's sampled competitor under
    ``nth_price_auction(n=2, num_winners=1)`` (synthetic_kw_helpers.py:116-180), i.e. it wins iff
    its bid strictly exceeds every rival and the competitor, and pays the largest of them.  Ties
    at the top win nothing.  All bidders of a world see the same keyword volumes and competitor
    draws (``env_group`` in the C ABI); clicks, conversions, revenues, budgets, rewards and
    episode state are per bidder.

    Rows of the wrapped VectorBiddingSimulation are ``world * A + agent``.  ``step`` takes bids
    ``[worlds, A, K]`` (device) and optional budgets ``[worlds, A]`` and returns observations with
    a leading ``[worlds, A]``."""

    def __init__(self, num_agents: int, num_worlds: int, **env_kwargs):
        self.num_agents, self.num_worlds = int(num_agents), int(num_worlds)
        env_kwargs["env_group"] = self.num_agents
        env_kwargs.setdefault("shared_keywords", True)
        self.vec = VectorBiddingSimulation(self.num_agents * self.num_worlds, **env_kwargs)

    @staticmethod
    def rival_floor_cents(bids: torch.Tensor) -> torch.Tensor:
        """[worlds, A, K] dollars -> int32 cents of the highest RIVAL bid per bidder (bids are
        canonicalised like the env does, ``round(max(bid, 0.01), 2)``, gymnasium_kw_env.py:215).
        Tensor restatement of what the kernels compute per unit (``unit_floor`` in csrc/adc_step.cu);
        ``step`` does not call it -- pass its result as ``floor_cents`` to ``vec.step`` to A/B the two."""
        cents = torch.round(torch.clamp(bids.to(torch.float64), min=0.01) * 100.0).to(torch.int32)
        first = cents.amax(dim=1, keepdim=True)
        is_top = cents == first
        lowest = torch.iinfo(torch.int32).min
        second = torch.where(is_top, torch.full_like(cents, lowest), cents).amax(dim=1, keepdim=True)
        # the unique top bidder faces the runner-up; everybody else (and tied leaders) faces the top bid
        unique_top = is_top & (is_top.sum(dim=1, keepdim=True) == 1)
        return torch.where(unique_top, second, first).contiguous()

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, info = self.vec.reset(seed=seed, options=options)
        return self._split(obs), info

    def _split(self, obs):
        W, A = self.num_worlds, self.num_agents
        return {k: v.view(W, A, *v.shape[1:]) for k, v in obs.items()}

    def step(self, bids: torch.Tensor, budget: Optional[torch.Tensor] = None, *, force_serial: bool = False):
        W, A, K = self.num_worlds, self.num_agents, self.vec.num_keywords
        assert tuple(bids.shape) == (W, A, K)
        # the kernels find every bidder's highest rival themselves (adc_step_args.env_group without a
        # floor_cents table): one pass over the A bid rows of the world per unit
        action = {"keyword_bids": bids.reshape(W * A, K).contiguous()}
        if budget is not None:
            action["budget"] = budget.reshape(W * A).contiguous()
        obs, reward, term, trunc, info = self.vec.step(action, force_serial=force_serial)
        return self._split(obs), reward.view(W, A), term.view(W, A), trunc.view(W, A), info


def make_multi_flat(num_agents: int, **env_kwargs) -> MultiAgentBiddingSimulation:
    """``adcraft/multi_agent/env.py:8`` -- one world of ``num_agents`` independent bidders."""
    return MultiAgentBiddingSimulation(num_agents, 1, **env_kwargs)
