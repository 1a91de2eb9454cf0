"""N > 1 host logic on CPU: two gloo ranks shard the envs, accumulate metrics locally and
all-reduce the metric vector; the result equals the single-process computation."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_envs, K, out):
    import torch.distributed as dist
    from adcraft_b200 import metrics as m
    from adcraft_b200.sharding import env_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = env_range(total_envs, rank, world)
    rng = np.random.default_rng(123)  # same stream on every rank: global arrays, local slices
    acc = m.MetricAccumulator(hi - lo, K, "cpu")
    for t in range(5):
        rev = rng.random((total_envs, K)); cost = rng.random((total_envs, K)); ideal = rng.random((total_envs, K))
        obs = {"revenue": torch.tensor(rev[lo:hi]), "cost": torch.tensor(cost[lo:hi])}
        acc.update(obs, torch.tensor((rev - cost)[lo:hi].sum(1)), ideal=torch.tensor(ideal[lo:hi]))
    vec = m.reduce_metrics(acc.summary_vector())
    # optional observation gather to the learner rank: rank order = global env order
    from adcraft_b200.sharding import gather_observations
    flat = torch.arange(lo * 6, hi * 6, dtype=torch.float32).view(hi - lo, 6)
    whole = gather_observations(flat, dst=0)
    try:
        gather_observations(flat, dst=0, max_bytes=16)
        refused = False
    except ValueError:
        refused = True
    if rank == 0:
        ok = torch.equal(whole, torch.arange(total_envs * 6, dtype=torch.float32).view(total_envs, 6)) and refused
        out.put((vec.numpy().copy(), bool(ok)))
    else:
        assert whole is None and refused
    dist.destroy_process_group()


def test_two_rank_metric_reduction_matches_single_process():
    import torch.multiprocessing as mp
    from adcraft_b200 import metrics as m
    total_envs, K, world = 10, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_envs, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, gathered_ok = q.get(timeout=120)
    assert gathered_ok, "gather_observations did not return the rows in global env order"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(123)
    acc = m.MetricAccumulator(total_envs, K, "cpu")
    for t in range(5):
        rev = rng.random((total_envs, K)); cost = rng.random((total_envs, K)); ideal = rng.random((total_envs, K))
        acc.update({"revenue": torch.tensor(rev), "cost": torch.tensor(cost)}, torch.tensor((rev - cost).sum(1)),
                   ideal=torch.tensor(ideal))
    np.testing.assert_allclose(got, acc.summary_vector().numpy(), rtol=1e-12)
