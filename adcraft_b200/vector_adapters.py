"""Vector-env front ends for RL libraries over ``VectorBiddingSimulation``.

The reference trains with RLlib / Stable-Baselines3 on ``FlatArrayWrapper(BiddingSimulation())``
(``adcraft/wrappers/flat_array.py:21-24,44-87``, ``RL/train_agent.ipynb`` cells 8, 10): flat
``[5K+2]`` observations and ``[K+1]`` actions in gymnasium's sorted-key order, one Python env per
rollout worker.  These classes present E envs of ONE device launch through the three vector
protocols those libraries consume.  None of the libraries is imported (they are not installed
here); the classes are duck-typed to the protocol each library calls:

* ``GymnasiumVectorAdapter``  gymnasium.vector.VectorEnv: ``reset(seed, options)``,
  ``step(actions) -> obs, rewards, terminations, truncations, infos`` with same-step autoreset and
  ``infos["final_observation"] / ["_final_observation"]``;
* ``SB3VecEnvAdapter``        stable_baselines3 VecEnv: ``reset()``, ``step_async`` /
  ``step_wait`` -> ``obs, rewards, dones, infos`` with ``terminal_observation`` and
  ``TimeLimit.truncated`` per finished env, ``get_attr`` / ``set_attr`` / ``env_method`` / ``seed``;
* ``RLlibVectorAdapter``      ray.rllib VectorEnv: ``vector_reset``, ``reset_at``, ``vector_step``,
  no autoreset (RLlib resets finished sub-envs itself).

Everything a learner on the GPU needs stays on the device (``flat=True, to_numpy=False``); with
``to_numpy=True`` one contiguous observation block comes back per step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .spaces import Box, get_action_space, get_observation_space
from .vector_env import VectorBiddingSimulation
from .wrappers import OBS_KEYS_SORTED, flat_observations, unflatten_actions


def _flat_spaces(K: int):
    obs = Box(-np.inf, np.inf, shape=(5 * K + 2,), dtype=np.float32)
    act = Box(0.01, np.inf, shape=(K + 1,), dtype=np.float32)
    return obs, act


class _Base:
    def __init__(self, env: VectorBiddingSimulation, flat: bool = True, to_numpy: bool = True):
        self.env = env
        self.num_envs = env.num_envs
        self.flat, self.to_numpy = flat, to_numpy
        K = env.num_keywords
        if flat:
            self.single_observation_space, self.single_action_space = _flat_spaces(K)
        else:
            self.single_observation_space = get_observation_space(K, env.budget)
            self.single_action_space = get_action_space(K)
        self.observation_space, self.action_space = self.single_observation_space, self.single_action_space

    # -- conversions ---------------------------------------------------------------------------
    def _action(self, actions) -> Dict[str, torch.Tensor]:
        dev = self.env.device
        if self.flat:
            a = torch.as_tensor(actions, device=dev)
            a = a.reshape(self.num_envs, -1)
            act = unflatten_actions(a)
            return {"keyword_bids": act["keyword_bids"].contiguous(), "budget": act["budget"].contiguous()}
        out = {"keyword_bids": torch.as_tensor(actions["keyword_bids"], device=dev).reshape(self.num_envs, -1)}
        if "budget" in actions:
            out["budget"] = torch.as_tensor(actions["budget"], device=dev).reshape(self.num_envs)
        return out

    def _obs(self, obs: Dict[str, torch.Tensor]):
        if self.flat:
            o = flat_observations(obs)
            return o.cpu().numpy() if self.to_numpy else o
        if self.to_numpy:
            return {k: obs[k].cpu().numpy() for k in OBS_KEYS_SORTED}
        return {k: obs[k] for k in OBS_KEYS_SORTED}

    def _host(self, t: torch.Tensor):
        return t.cpu().numpy() if self.to_numpy else t

    @staticmethod
    def _row(obs, i):
        return {k: v[i] for k, v in obs.items()} if isinstance(obs, dict) else obs[i]

    @staticmethod
    def _zero_rows(obs, mask):
        """The observation after ``reset`` is all zeros (env:331-343)."""
        if isinstance(obs, dict):
            return {k: _Base._zero_rows(v, mask) for k, v in obs.items()}
        out = obs.copy() if isinstance(obs, np.ndarray) else obs.clone()
        out[mask] = 0
        return out

    def close(self):
        self.env.close()


class GymnasiumVectorAdapter(_Base):
    """gymnasium.vector.VectorEnv protocol, same-step autoreset (gymnasium 0.26-0.29): a finished
    env's returned observation is already the reset one and its last observation travels in
    ``infos["final_observation"]`` (object array, ``None`` elsewhere) under the mask
    ``infos["_final_observation"]``."""

    def __init__(self, env: VectorBiddingSimulation, flat: bool = True, to_numpy: bool = True):
        if not env.autoreset:
            raise ValueError("GymnasiumVectorAdapter needs a VectorBiddingSimulation with autoreset=True")
        super().__init__(env, flat, to_numpy)
        self.is_vector_env = True

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self._obs(obs), info

    def step(self, actions):
        obs, reward, term, trunc, info = self.env.step(self._action(actions))
        o = self._obs(obs)
        done = term | trunc
        infos = dict(info)
        if bool(done.any()):
            mask = done.cpu().numpy() if self.to_numpy else done
            final = np.full(self.num_envs, None, dtype=object)
            for i in np.flatnonzero(done.cpu().numpy()):
                final[i] = self._row(o, int(i))
            infos["final_observation"] = final
            infos["_final_observation"] = done.cpu().numpy()
            o = self._zero_rows(o, mask)
        return o, self._host(reward), self._host(term), self._host(trunc), infos


class SB3VecEnvAdapter(_Base):
    """stable_baselines3.common.vec_env.VecEnv protocol on flat observations / actions."""

    def __init__(self, env: VectorBiddingSimulation, to_numpy: bool = True):
        if not env.autoreset:
            raise ValueError("SB3VecEnvAdapter needs a VectorBiddingSimulation with autoreset=True")
        super().__init__(env, flat=True, to_numpy=to_numpy)
        self.render_mode = env.render_mode
        self._pending = None
        self._seed: Optional[int] = None

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        self._seed = seed
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def reset(self):
        obs, _ = self.env.reset(seed=self._seed)
        self._seed = None
        return self._obs(obs)

    def step_async(self, actions) -> None:
        self._pending = self.env.step(self._action(actions))  # enqueued on the stream, not synchronised

    def step_wait(self):
        obs, reward, term, trunc, _ = self._pending
        self._pending = None
        o = self._obs(obs)
        done = term | trunc
        infos: List[dict] = [{} for _ in range(self.num_envs)]
        if bool(done.any()):
            trunc_h = trunc.cpu().numpy()
            term_h = term.cpu().numpy()
            for i in np.flatnonzero(done.cpu().numpy()):
                infos[i]["terminal_observation"] = self._row(o, int(i)).copy() if self.to_numpy else self._row(o, int(i)).clone()
                infos[i]["TimeLimit.truncated"] = bool(trunc_h[i] and not term_h[i])
            o = self._zero_rows(o, done.cpu().numpy() if self.to_numpy else done)
        return o, self._host(reward), self._host(done), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_attr(self, attr_name: str, indices=None) -> list:
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self.env, attr_name)] * n

    def set_attr(self, attr_name: str, value, indices=None) -> None:
        setattr(self.env, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> list:
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self.env, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [False] * n

    def _indices(self, indices) -> Sequence[int]:
        return [indices] if isinstance(indices, int) else list(indices)


class RLlibVectorAdapter(_Base):
    """ray.rllib.env.VectorEnv protocol: lists of per-env observations, ``reset_at`` for the
    sub-envs RLlib decides to reset (the wrapped env must be built with ``autoreset=False``)."""

    def __init__(self, env: VectorBiddingSimulation, flat: bool = True):
        if env.autoreset:
            raise ValueError("RLlibVectorAdapter needs a VectorBiddingSimulation with autoreset=False")
        super().__init__(env, flat, to_numpy=True)

    def vector_reset(self, *, seeds: Optional[List[int]] = None, options: Optional[List[dict]] = None):
        obs, info = self.env.reset(seed=None if not seeds else seeds[0],
                                   options=None if not options else options[0])
        o = self._obs(obs)
        return [self._row(o, i) for i in range(self.num_envs)], [info] * self.num_envs

    def reset_at(self, index: Optional[int] = None, *, seed: Optional[int] = None, options: Optional[dict] = None):
        mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.env.device)
        mask[0 if index is None else index] = 1
        obs = self.env.reset_envs(mask)
        o = self._obs(obs)
        return self._row(o, 0 if index is None else index), {}

    def vector_step(self, actions):
        act = np.stack([np.asarray(a) for a in actions]) if self.flat else {
            k: np.stack([np.asarray(a[k]) for a in actions]) for k in actions[0]}
        obs, reward, term, trunc, info = self.env.step(self._action(act))
        o = self._obs(obs)
        return ([self._row(o, i) for i in range(self.num_envs)], list(reward.cpu().numpy()),
                list(term.cpu().numpy()), list(trunc.cpu().numpy()), [dict(info) for _ in range(self.num_envs)])

    def get_sub_environments(self):
        return []
