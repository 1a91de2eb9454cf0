import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import make_implicit_table
from adcraft_b200 import keywords as kwm
from adcraft_b200.vector_env import VectorBiddingSimulation
from oracle import oracle as orc
orc.build()
rng = np.random.default_rng(1)
K, E = 5, 4
table = make_implicit_table(rng, K, 128)
env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, device="cuda", seed=4242, budget=1e5, obs_dtype=torch.float64)
env.reset()
ob = orc.BatchOracle(table.kind, E, K, {n: getattr(table, n) for n in kwm.PARAM_NAMES}, seed=4242, budget=1e5)
bids = np.round(rng.uniform(0.05, 1.6, (E, K)), 2)
obs, reward, *_ = env.step({"keyword_bids": torch.from_numpy(bids).cuda()})
ref = ob.step(bids, n_threads=1)
for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"), ("sellside_conversions", "conversions"), ("cost","cost"),("revenue","revenue")):
    g = obs[a].cpu().numpy()
    print(a, "\n", g, "\n", ref[b])
