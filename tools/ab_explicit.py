"""A/B of the explicit-keyword kernels on the reference's default env shape: flattened (n_lanes=0) against one
thread per unit (n_lanes=1), kernel time from CUDA events around a graph-free loop of library calls."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from adcraft_b200.vector_env import VectorBiddingSimulation
from adcraft_b200 import keywords as kwm
for E in (4096, 65536):
    for K in (10, 100):
        table = kwm.sample_random_keywords(K, np.random.default_rng(0))
        bids = torch.from_numpy(np.round(np.random.default_rng(1).uniform(0.01, 3.0, (E, K)), 2).astype(np.float32)).cuda()
        res = {}
        for nl in (0, 1):
            env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e6, device="cuda", seed=3, n_lanes=nl)
            env.reset()
            act = {"keyword_bids": bids}
            for _ in range(5):
                obs = env.step(act)[0]
            torch.cuda.synchronize()
            n = 30
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                obs = env.step(act)[0]
            t.record()
            torch.cuda.synchronize()
            res[nl] = (s.elapsed_time(t) / n, int(obs["impressions"].sum()), int(obs["buyside_clicks"].sum()))
        slots = float(np.mean(table.vol_mean)) + 23
        print(f"E={E} K={K}: flattened {res[0][0]*1e3:.1f} us, thread-per-unit {res[1][0]*1e3:.1f} us per step; "
              f"{E*K/(res[0][0]*1e-3):.3g} units/s, {E*K*slots/(res[0][0]*1e-3):.3g} slots/s (auctions + phantom slots); same results: {res[0][1:] == res[1][1:]}")
