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


def make_multi_flat(num_agents: int, **env_kwargs) -> MultiAgentBiddingSimulation:
    """``adcraft/multi_agent/env.py:8`` -- one world of ``num_agents`` independent bidders."""
    return MultiAgentBiddingSimulation(num_agents, 1, **env_kwargs)
