"""TEST / BENCH INFRASTRUCTURE ONLY -- time the reference's own BiddingSimulation.step on host cores.

The UNMODIFIED reference Python (staged by ``ref_harness.stage_reference`` into the git-ignored
``baseline/_ref``, or read from /root/reference where that exists) runs one env per process:
``bidding_sim_creator`` with the experiment keyword config (experiment_configs.py:15-27), fixed
bids, scalar budget -- the CPU arm of bench.py (``--impl reference``) and its ``cpu_baseline`` leg.
The reference is single-threaded Python under the GIL, so "all host cores" means one independent
env per core.  The nine helpers of its PyO3 module (src/lib.rs) come from ``ref_harness.RustShim``
(numpy restatement): there is no Rust toolchain in the image, and the report says so.
"""
from __future__ import annotations

import os
import tempfile
import time
from typing import Dict, Optional


def reference_root() -> Optional[str]:
    from . import ref_harness as rh
    for root in (rh.STAGED_ROOT, rh.REFERENCE_ROOT):
        if rh.reference_available(root):
            return root
    return None


def _worker(job) -> Dict[str, float]:
    rank, root, K, vol, cvr, bid, budget, max_days, steps, warmup, seed, drift = job
    import numpy as np
    from . import ref_harness as rh
    ref = rh.load_reference(root)
    tmp = tempfile.mkdtemp(prefix="adcraft_ref_")
    env = ref["env"].bidding_sim_creator(dict(
        keyword_config=rh.experiment_keyword_config(vol, cvr, tmp), num_keywords=K, max_days=max_days,
        updater_mask=[True] * K if drift else None))
    env.reset(seed=seed)
    bids = np.full(K, bid)
    action = {"keyword_bids": bids, "budget": budget}
    n_auctions = 0

    def one():
        nonlocal n_auctions
        obs, _r, term, trunc, _info = env.step(action)
        if term or trunc:
            env.reset()
        return obs

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        obs = one()
    dt = time.perf_counter() - t0
    return {"seconds": dt, "units": float(K * steps), "impressions_last_step": float(np.sum(obs["impressions"]))}


def time_reference(K: int, vol: int, cvr: float, bid: float, budget: float, max_days: int, steps: int,
                   warmup: int, procs: int, drift: bool = False, seed: int = 5,
                   timeout_s: float = 600.0) -> Optional[Dict[str, float]]:
    """Run ``procs`` independent reference envs (one per process) for ``steps`` timed steps each.
    Returns units/s over all processes (units of all / slowest process' time), or None when no
    reference tree is available.  Workers are plain ``python -m oracle.ref_bench --worker`` child
    processes (the GPU arm's parent holds a CUDA context: nothing is forked from it)."""
    root = reference_root()
    if root is None:
        return None
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1",
               CUDA_VISIBLE_DEVICES="")
    job = json.dumps([root, K, vol, cvr, bid, budget, max_days, steps, warmup, seed, drift])
    t0 = time.perf_counter()
    children = [subprocess.Popen([sys.executable, "-m", "oracle.ref_bench", "--worker", str(r), job], cwd=repo,
                                 env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                for r in range(procs)]
    res = []
    try:
        for c in children:
            out, err = c.communicate(timeout=max(1.0, timeout_s - (time.perf_counter() - t0)))
            if c.returncode != 0:
                raise RuntimeError(f"reference worker failed: {err[-2000:]}")
            res.append(json.loads(out.strip().splitlines()[-1]))
    finally:
        for c in children:
            if c.poll() is None:
                c.kill()
    wall = time.perf_counter() - t0
    slowest = max(r["seconds"] for r in res)
    units = sum(r["units"] for r in res)
    return {"units_per_s": units / slowest, "s_per_step": slowest / steps, "procs": procs, "wall_s": wall,
            "units_per_s_per_core": units / slowest / procs, "root": root,
            "impressions_per_step": sum(r["impressions_last_step"] for r in res) / procs}


if __name__ == "__main__":
    import json
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "--worker":
        print(json.dumps(_worker((int(sys.argv[2]), *json.loads(sys.argv[3])))))
