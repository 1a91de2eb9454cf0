import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import bench
from adcraft_b200.vector_env import VectorBiddingSimulation
from adcraft_b200 import _capi
import os
E, K = int(os.environ.get("DBG_E", 4096)), int(os.environ.get("DBG_K", 100))
bench.K_KW, bench.MEAN_VOLUME, bench.CVR = K, int(os.environ.get("DBG_V", 128)), float(os.environ.get("DBG_CVR", 0.8))
table = bench.workload_table()
env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, device="cuda", seed=1234, budget=1000.0)
env.reset()
bids = torch.full((E, K), 0.75, device="cuda", dtype=torch.float32)
for _ in range(4):
    env.step({"keyword_bids": bids})
torch.cuda.synchronize()
lib = _capi.load()
buf = np.zeros(8192 * 12, dtype=np.uint64)
lib.adc_debug_read(C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes))
t = buf.reshape(-1, 12)[:E].astype(np.int64)
t0 = t[:, 0].min()
print("start spread us", (t[:, 0].max() - t0) / 1e3, "end max us", (t[:, 3].max() - t0) / 1e3)
for name, a, b in (("expand", 0, 1), ("walk", 1, 2), ("commit", 2, 3), ("total", 0, 3)):
    d = (t[:, b] - t[:, a]) / 1e3
    print(name, "mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f us" % (d.mean(), np.percentile(d, 50), np.percentile(d, 90), np.percentile(d, 99), d.max()))
end = (t[:, 3] - t0) / 1e3
print("end percentiles", np.percentile(end, [10, 50, 90, 99, 100]))
print("serial count", int(env._scratch["serial_count"].max()), "hint sum", int(env._scratch["serial_hint"].sum()))

names = ["skip", "none", "all", "gen", "lanes_scanned", "direct", "clicks_scanned"]
for i, n in enumerate(names):
    c = t[:, 4 + i]
    print(n, "mean %.1f p50 %d p99 %d max %d" % (c.mean(), np.percentile(c, 50), np.percentile(c, 99), c.max()))
walk = (t[:, 2] - t[:, 1]) / 1e3
slow = np.argsort(-walk)[:8]
for i in slow:
    print("slow env", t[i, 11], "walk %.0f us" % walk[i], dict(zip(names, t[i, 4:11])))
print("corr walk vs gen", np.corrcoef(walk, t[:, 7])[0, 1], "vs lanes_scanned", np.corrcoef(walk, t[:, 8])[0, 1], "vs direct", np.corrcoef(walk, t[:, 9])[0, 1])
