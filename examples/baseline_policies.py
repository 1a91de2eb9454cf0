#!/usr/bin/env python
"""The reference's two baseline bidders driving E environments at once.

    python examples/baseline_policies.py --envs 1024 --keywords 100 --days 60 --policy interpolation

Mirrors the loop of run_heatmap_experiments.ipynb (cell 3): observe, update the agent's caches, sample
the next action -- with the vectorised agents of adcraft_b200.baselines instead of one Python agent
per env, the env stepping on the GPU and the AKNCP / NCP metrics accumulated on the device.
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
from adcraft_b200.baselines import VectorNaiveInterpolationStrategy, VectorNaiveZeroMarginStrategy  # noqa: E402
from adcraft_b200.vector_env import VectorBiddingSimulation  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--keywords", type=int, default=100)
    ap.add_argument("--days", type=int, default=60)
    ap.add_argument("--volume", type=int, default=128)
    ap.add_argument("--cvr", type=float, default=0.8)
    ap.add_argument("--policy", choices=["interpolation", "zero_margin"], default="interpolation")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    E, K = args.envs, args.keywords
    env = VectorBiddingSimulation(E, keyword_config={"mean_volume": args.volume, "conversion_rate": args.cvr},
                                  num_keywords=K, max_days=args.days, device=dev, seed=args.seed,
                                  obs_dtype=torch.float64)
    env.reset(seed=args.seed)
    pol = (VectorNaiveInterpolationStrategy(E, K, device=dev, seed=args.seed) if args.policy == "interpolation"
           else VectorNaiveZeroMarginStrategy(E, K, device=dev, seed=args.seed))
    ideal = M.ideal_profit(env)["ideal"]  # [1, K]: the keyword set is shared by all envs
    acc = M.MetricAccumulator(E, K, dev)
    if args.policy == "interpolation":
        action = pol.sample_action()
    else:  # the zero-margin agent ramps up from its first observation (interpolated_expectations.py:483-515)
        action = {"keyword_bids": torch.full((E, K), 0.01, dtype=torch.float64, device=dev),
                  "budget": torch.full((E,), 1000.0, dtype=torch.float64, device=dev)}
    t0 = time.perf_counter()
    total_reward = torch.zeros(E, dtype=torch.float64, device=dev)
    for _ in range(args.days):
        obs, reward, term, trunc, _ = env.step({"keyword_bids": action["keyword_bids"], "budget": action["budget"]})
        total_reward += reward
        acc.update(obs, reward, ideal=ideal, done=term)
        pol.update_all_caches(action, obs)
        action = pol.sample_action()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    out = {"policy": args.policy, "envs": E, "keywords": K, "days": args.days,
           "mean_episode_profit": float(total_reward.mean()), "seconds": dt,
           "env_steps_per_s": E * args.days / dt, "keyword_auction_steps_per_s": E * K * args.days / dt}
    out.update(M.summarize(M.reduce_metrics(acc.summary_vector())))
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
