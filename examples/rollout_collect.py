#!/usr/bin/env python
"""PPO-style rollout collection on the batched simulator (BASELINE.json configs[4] in miniature).

    python examples/rollout_collect.py --envs 4096 --keywords 100 --steps 60
    python -m torch.distributed.run --nproc-per-node 8 examples/rollout_collect.py --envs 131072 --keywords 10000

Each rank owns `--envs` environments (global env ids = rank * envs + i, so the draws do not depend
on the GPU count), runs a small MLP policy replica on the flat observations ([5K+2], the
reference's FlatArrayWrapper layout) to produce [budget | bids], steps the simulator, stores the
rollout on the device, accumulates per-keyword profits against the ideal-profit estimate
(experiment_metrics.py:20-61) and all-reduces the AKNCP / NCP summary over NCCL once per episode.
The policy and the optimiser are plain torch: they are not the hot path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from adcraft_b200 import metrics as M  # noqa: E402
from adcraft_b200.sharding import init_distributed  # noqa: E402
from adcraft_b200.vector_env import VectorBiddingSimulation  # noqa: E402
from adcraft_b200.wrappers import flat_observations  # noqa: E402


class Policy(torch.nn.Module):
    def __init__(self, K: int, hidden: int = 64):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(5 * K + 2, hidden), torch.nn.Tanh(),
                                        torch.nn.Linear(hidden, hidden), torch.nn.Tanh())
        self.mu = torch.nn.Linear(hidden, K)
        self.log_std = torch.nn.Parameter(torch.full((K,), -1.5))
        self.value = torch.nn.Linear(hidden, 1)

    def forward(self, obs):
        h = self.body(torch.log1p(obs.clamp(min=0)) * 0.2)
        mu = 0.75 + 0.5 * torch.tanh(self.mu(h))
        return mu, self.log_std.exp(), self.value(h).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--keywords", type=int, default=100)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--volume", type=int, default=128)
    ap.add_argument("--cvr", type=float, default=0.8)
    ap.add_argument("--seed", type=int, default=0x5EED)
    args = ap.parse_args()
    rank, world = init_distributed("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    E, K, T = args.envs, args.keywords, args.steps
    env = VectorBiddingSimulation(
        E, keyword_config={"mean_volume": args.volume, "conversion_rate": args.cvr}, num_keywords=K,
        budget=100000.0, max_days=T, device=dev, seed=args.seed, env_base=rank * E)
    obs, _ = env.reset(seed=5)
    torch.manual_seed(1234)  # same policy replica on every rank
    policy = Policy(K).to(dev)
    ideal = M.ideal_profit(env)["ideal"]  # [1, K]: the keyword set is shared by all envs
    acc = M.MetricAccumulator(E, K, dev)
    buf_obs = torch.empty(T, E, 5 * K + 2, device=dev) if E * K * T < 2e9 else None
    buf_act = torch.empty(T, E, K, device=dev) if buf_obs is not None else None
    buf_rew = torch.empty(T, E, device=dev)
    budget = torch.full((E,), 100000.0, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(T):
            flat = flat_observations(obs)
            mu, std, _v = policy(flat)
            bids = (mu + std * torch.randn_like(mu)).clamp(min=0.01)
            obs, reward, term, trunc, _ = env.step({"keyword_bids": bids, "budget": budget})
            acc.update(obs, reward, ideal=ideal, done=term)
            if buf_obs is not None:
                buf_obs[i], buf_act[i] = flat, bids
            buf_rew[i] = reward
    vec = M.reduce_metrics(acc.summary_vector())  # the only collective: 8 doubles over NCCL
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if rank == 0:
        out = M.summarize(vec)
        out.update(ranks=world, envs_per_rank=E, keywords=K, steps=T,
                   units_per_s=world * E * K * T / dt, mean_step_reward=float(buf_rew.mean()))
        print(json.dumps(out))


if __name__ == "__main__":
    main()
