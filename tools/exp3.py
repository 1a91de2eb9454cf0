import sys, time, torch, ctypes as C
sys.path.insert(0, ".")
import bench
from adcraft_b200.vector_env import VectorBiddingSimulation
E, K = 4096, 100
table = bench.workload_table()
env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e7, device="cuda", seed=1, episode_profit=True)
env.reset()
bids = torch.full((E, K), 0.75, dtype=torch.float32).pin_memory()
bids_dev = bids.cuda()
def run(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("records, sync per step      : %.1f us" % run(lambda: env.step_host_records(bids)))
orig = torch.cuda.Stream.synchronize
torch.cuda.Stream.synchronize = lambda self: None
print("records, no sync (queue)    : %.1f us" % run(lambda: env.step_host_records(bids)))
# same with device bids: patch the pinned check
import adcraft_b200.vector_env as ve
class Fake:
    pass
def rec_dev():
    a = env._fill_args(bids_dev, None, False)
    out = a.out
    saved = (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated)
    (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = env._rec_ptrs
    out.unit_records = env._rec_host.data_ptr()
    stream = torch.cuda.current_stream(env.device)
    env._call(env._lib.adc_step_philox, C.byref(a), C.c_void_p(stream.cuda_stream))
    (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = saved
    out.unit_records = None
    env._step_count += 1; env._calls += 1
print("records, device bids (queue): %.1f us" % run(rec_dev))
def dev_only():
    env.step({"keyword_bids": bids_dev})
print("device step (queue)         : %.1f us" % run(dev_only))
def dev_hostbids():
    a = env._fill_args(bids, None, False)
    stream = torch.cuda.current_stream(env.device)
    env._call(env._lib.adc_step_philox, C.byref(a), C.c_void_p(stream.cuda_stream))
    env._step_count += 1; env._calls += 1
print("device outputs, host bids (queue): %.1f us" % run(dev_hostbids))
