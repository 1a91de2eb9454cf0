"""Where a step_host(mode=MODE) call spends its time on the host: Python / ctypes before the launch,
the launch, and the wait for the kernels (C2 shape)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from adcraft_b200.vector_env import VectorBiddingSimulation
MODE = sys.argv[1] if len(sys.argv) > 1 else "zero_copy"
E, K = 4096, 100
table = bench.workload_table()
env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e7, device="cuda", seed=1, episode_profit=True)
env.reset()
bids = torch.full((E, K), 0.75, dtype=torch.float32).pin_memory()
for _ in range(20):
    env.step_host(bids, mode=MODE)
n = 200
t0 = time.perf_counter()
for _ in range(n):
    env.step_host(bids, mode=MODE)
t1 = time.perf_counter()
print("step_host: %.1f us per call" % ((t1 - t0) / n * 1e6))
# the same without waiting: host-side cost of a call
orig = torch.cuda.Stream.synchronize
torch.cuda.Stream.synchronize = lambda self: None
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    env.step_host(bids, mode=MODE)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
torch.cuda.Stream.synchronize = orig
print("host side only: %.1f us per call (queue drained after %.1f us per call)" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
import cProfile, pstats
torch.cuda.Stream.synchronize = lambda self: None
pr = cProfile.Profile(); pr.enable()
for _ in range(n):
    env.step_host(bids, mode=MODE)
pr.disable(); torch.cuda.synchronize()
torch.cuda.Stream.synchronize = orig
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
