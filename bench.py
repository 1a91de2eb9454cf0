#!/usr/bin/env python
"""Benchmark of the BiddingSimulation.step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric: keyword-auction-steps/s = envs x keywords x env-steps / second.
Workload at N=1 (BASELINE.json configs[1]): 4096 vectorised envs x 100 synthetic dense implicit
keywords (experiment_configs.py:15-27 parameters, keyword set drawn once from default_rng(5) and
shared by all envs), fixed bid 0.75, budget 100000, 60-day episodes with auto-reset.  With N > 1
(torchrun, one rank per GPU) every rank owns 4096 envs (weak scaling); env ids are global, so
results do not depend on N.

One JSON line is printed by rank 0 (see the round contract): `value` is device-timed with the
inputs resident in HBM, `e2e` is the same metric through VectorBiddingSimulation.step_host with
pinned HOST buffers (H2D + D2H inside the timed region), `roofline` compares the dominant kernel
with the measured HBM peak, `cpu_baseline` is the reference's own Python (staged copy under
baseline/_ref, one env per host core) timed on this box, with the C oracle port beside it.
`--impl reference` times only that CPU reference on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version
    banner to stdout when the box sets NCCL_DEBUG), so file descriptor 1 is pointed at stderr for
    the whole process and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


_OUT = None


def emit(obj) -> None:
    (_OUT or sys.stdout).write(json.dumps(obj) + "\n")
    (_OUT or sys.stdout).flush()


METRIC = "keyword_auction_steps_per_sec"
UNIT = "keyword-auction-steps/s"
K_KW, E_ENVS, BID, BUDGET, MAX_DAYS = 100, 4096, 0.75, 100000.0, 60
SEED = 0x5EED
B_UNIT = 24  # read bid 4 B + write 3 x int32 + 2 x f32 per (env, keyword, step)
HOT_KERNEL = "adc_flat2_implicit_kernel"
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_flat2_c2_ncu.json")


def ncu_capture_of_hot_kernel():
    """DRAM traffic and warp-instruction count of one hot-kernel launch on the C2 workload, read
    from the committed summary of the `ncu --set full` capture (profiles/r02_flat2_c2_ncu.json,
    written by tools/ncu_summary.py from the .ncu-rep).  Returns None -- loudly -- when the capture
    does not describe what this run launches (other workload shape, or the kernel symbol is not in
    the loaded library): stale numbers are not reported."""
    try:
        d = json.load(open(NCU_SUMMARY))
    except (OSError, ValueError) as exc:
        print(f"bench.py: no ncu summary ({exc}); roofline.traffic = null", file=sys.stderr)
        return None
    from adcraft_b200 import build as b
    lib_has = HOT_KERNEL.encode() in open(b.LIB_PATH, "rb").read()
    same = (d.get("kernel", "").find(HOT_KERNEL) >= 0 and d.get("envs") == E_ENVS and d.get("keywords") == K_KW
            and d.get("mean_volume") == MEAN_VOLUME and not DRIFT)
    if not (lib_has and same):
        print(f"bench.py: the committed ncu capture ({d.get('kernel')}, {d.get('envs')} x {d.get('keywords')}) does "
              f"not describe this run ({HOT_KERNEL} in library: {lib_has}; {E_ENVS} x {K_KW}); roofline.traffic = null",
              file=sys.stderr)
        return None
    return d


MEAN_VOLUME, CVR, DRIFT = 128, 0.8, False


def workload_table():
    from adcraft_b200 import keywords as kwm
    rng = np.random.default_rng(5)
    return kwm.sample_implicit_keywords_from_quantiles(
        K_KW, rng, {"mean_volume": MEAN_VOLUME, "conversion_rate": CVR})


def config_dict(n_gpus, extra=None):
    c = {
        "workload": "C2: BiddingSimulation, 100 synthetic dense implicit keywords x 4096 vectorised envs "
                    "per GPU, fixed bid 0.75, budget 100000, free-running Philox draws",
        "envs_per_gpu": E_ENVS, "keywords": K_KW, "mean_volume": MEAN_VOLUME, "conversion_rate": CVR,
        "non_stationary": DRIFT,
        "episode_days": MAX_DAYS, "parallelism": f"env-sharded x{n_gpus}",
        "l2": "flushed between timed steps (256 MiB memset outside the event pairs)",
    }
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------
# CPU port (oracle) timing: cpu_baseline leg and --impl reference
# --------------------------------------------------------------------------------------------
def time_cpu_port(steps: int, warmup: int, envs: int, threads: int):
    from oracle import oracle as orc
    from adcraft_b200 import keywords as kwm
    orc.build()
    table = workload_table()
    params = {n: getattr(table, n) for n in kwm.PARAM_NAMES}
    ob = orc.BatchOracle(orc.IMPLICIT, envs, K_KW, params, seed=SEED, budget=BUDGET, max_days=MAX_DAYS)
    bids = np.full((envs, K_KW), BID)
    for _ in range(warmup):
        ob.step(bids, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ob.step(bids, n_threads=threads)
    dt = time.perf_counter() - t0
    return envs * K_KW * steps / dt, dt / steps


def time_reference_python(steps: int, warmup: int, procs: int):
    """The reference's own Python BiddingSimulation.step (staged copy, see oracle/ref_bench.py), one
    env per host core.  Returns (units/s, s/step, info) or None when no reference tree travelled."""
    from oracle import ref_bench
    r = ref_bench.time_reference(K_KW, MEAN_VOLUME, CVR, BID, BUDGET, MAX_DAYS, steps, warmup, procs, drift=DRIFT)
    if r is None:
        return None
    return r["units_per_s"], r["s_per_step"], r


REF_NOTE = ("the reference's unmodified Python (gymnasium_kw_env / bidding_simulation / synthetic_kw_*), one "
            "single-threaded env per host core; the nine helpers of its PyO3 module come from a numpy "
            "restatement (oracle/ref_harness.RustShim): no Rust toolchain exists in the image")


def cpu_baseline_block(steps: int, warmup: int):
    """cpu_baseline of the JSON line: the reference itself when its staged copy is present
    (kind "reference"), else the C oracle port (kind "port")."""
    procs = os.cpu_count() or 1
    ref = time_reference_python(steps, warmup, procs)
    port_envs, port_steps = 256, 6
    port_v, _ = time_cpu_port(port_steps, 1, port_envs, procs)
    port = {"value": port_v, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{port_envs} of {E_ENVS} envs x {K_KW} keywords per step, {port_steps} steps, OpenMP over envs",
            "note": "C restatement of the reference algorithm (oracle/adcraft_oracle.c), for context"}
    if ref is None:
        return port, None
    v, s_per_step, info = ref
    return ({"value": v, "unit": UNIT, "cores": procs, "kind": "reference",
             "sample": f"{procs} independent envs (one per core) x {K_KW} keywords, {steps} timed steps each "
                       f"after {warmup} warm-up ({info['wall_s']:.1f} s wall incl. process start)",
             "units_per_s_per_core": info["units_per_s_per_core"], "note": REF_NOTE, "c_port": port}, s_per_step)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # each reference step takes ~0.35 s per env: bound the sample so the arm ends within minutes
    steps = min(args.steps, 40)
    warmup = min(args.warmup, 3)
    cpu, s_per_step = cpu_baseline_block(steps, warmup)
    if s_per_step is None:  # no staged reference: the C port is what ran
        threads = os.cpu_count() or 1
        cpu["value"], s_per_step = time_cpu_port(args.steps, args.warmup, 256, threads)
    val = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (numpy)",
        "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "timed_steps_per_env": steps,
    }
    emit(line)
    return 0


def bind_to_gpu_cpus(gpu_index: int):
    """One process per GPU: run on (and first-touch pinned host memory from) the CPUs NVML reports
    as local to that GPU, so the e2e leg's host buffers do not sit behind the inter-socket link.
    Best effort: a cpuset that excludes those CPUs leaves the affinity unchanged."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return {"cpus_before": len(before), "cpus_after": len(after), "first_cpu": after[0] if after else None}
    except Exception as exc:  # noqa: BLE001
        return {"error": type(exc).__name__}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled through NVML from a thread (a sample costs
    ~0.1 ms, so even the 4 ms timed region of a 20-step run holds several; nvidia-smi's loop mode
    cannot sample faster than every ~20 ms and needs ~0.1 s to come up)."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, torch, dev):
        import threading
        self.samples = []  # (perf_counter, sm_mhz, reason bits)
        self.t_mark = None
        self.h = None
        self.max_mhz = None
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(dev).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append((time.perf_counter(), sm, bits))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.0003)

    def mark(self):
        self.t_mark = time.perf_counter()

    def count_since_mark(self):
        return sum(1 for t, _, _ in self.samples if self.t_mark is not None and t >= self.t_mark)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self.h is None:
            return out
        self._stop.set()
        self.thread.join(timeout=2)
        rows = [(sm, bits) for t, sm, bits in self.samples if self.t_mark is None or t >= self.t_mark]
        if rows:
            reasons = set()
            for name, attr in self.REASONS:
                flag = getattr(self.nv, attr, 0)
                if any(bits & flag for _, bits in rows):
                    reasons.add(name)
            out = {"sm_mhz": float(np.median([sm for sm, _ in rows])), "sm_max_mhz": self.max_mhz,
                   "reasons": sorted(reasons), "samples": len(rows), "source": "NVML polled from a thread"}
        return out


# --------------------------------------------------------------------------------------------
# replay leg
# --------------------------------------------------------------------------------------------
def run_replay_leg(env, table, dev, torch, steps=20):
    """Time adc_step_replay on a synthetic pre-drawn tape of the same shape (resident in HBM,
    ~1.3 GB, i.e. 10x the L2) and report the HBM roofline of the replay kernel.  Algorithmic bytes
    per launch = sum over units of 4 V + 8 I + 8 B + 4 S (consumed stream entries) + 76 B/unit
    (record header and offset / volume and 4 CSR offsets, bid, outputs), with I, B, S read back from
    the step's own outputs."""
    from adcraft_b200.tape import DeviceTape
    E, K = env.num_envs, env.num_keywords
    g = torch.Generator(device=dev).manual_seed(1234)
    f64 = torch.float64
    vol_mean = torch.tensor(table.vol_mean, device=dev)
    vol_std = torch.tensor(table.vol_std, device=dev)
    V = torch.clamp(torch.round(vol_mean + vol_std * torch.randn(E, K, device=dev, dtype=f64, generator=g)), min=0)
    V = V.to(torch.int32)
    off = torch.zeros(E * K + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(V.reshape(-1).to(torch.int64), 0)
    n = int(off[-1])
    unit = torch.repeat_interleave(torch.arange(E * K, device=dev), V.reshape(-1).to(torch.int64))
    kw = unit % K
    loc = torch.tensor(table.p1, device=dev)[kw]
    scale = torch.tensor(table.p2, device=dev)[kw]
    u = torch.rand(n, device=dev, dtype=f64, generator=g)
    lap = torch.where(u >= 0.5, -torch.log(2.0 - 2.0 * u), torch.log(2.0 * u))
    comp = torch.round((loc + scale * lap).abs() * 100.0).to(torch.int32)
    rev = torch.clamp(torch.round((torch.tensor(table.rev_mean, device=dev)[kw] + torch.tensor(
        table.rev_std, device=dev)[kw] * torch.randn(n, device=dev, dtype=f64, generator=g)) * 100.0), min=1)
    loose = DeviceTape(V, off, comp, off, torch.rand(n, device=dev, dtype=f64, generator=g), off,
                       torch.rand(n, device=dev, dtype=f64, generator=g), off, rev.to(torch.int32))
    del unit, kw, loc, scale, u, lap
    bids = torch.full((E, K), BID, dtype=torch.float32, device=dev)
    action = {"keyword_bids": bids}
    # the synthetic streams are over-provisioned (V entries each); a recording holds exactly what
    # the step consumed, so replay once, cut every stream to its consumed length and pack
    obs0 = env.step_replay(action, loose)[0]
    keys = ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue")
    ref = {k: obs0[k].clone() for k in keys}
    tape = loose.trimmed(ref["impressions"], ref["buyside_clicks"], ref["sellside_conversions"]).pack()
    del loose
    csr_only = DeviceTape(**{k: v for k, v in tape.__dict__.items() if k not in ("packed", "packed_off")})

    def timed(tp):
        for _ in range(3):
            obs = env.step_replay(action, tp)[0]
        torch.cuda.synchronize(dev)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            starts[i].record()
            obs = env.step_replay(action, tp)[0]
            stops[i].record()
        torch.cuda.synchronize(dev)
        for k in keys:
            if not torch.equal(obs[k], ref[k]):
                raise SystemExit(f"bench.py: replay of the trimmed/packed tape changed {k}")
        return sum(s.elapsed_time(e) for s, e in zip(starts, stops)) / steps

    ms = timed(tape)
    ms_csr = timed(csr_only)
    I = int(ref["impressions"].sum()); B = int(ref["buyside_clicks"].sum()); S = int(ref["sellside_conversions"].sum())
    alg = 4 * n + 8 * I + 8 * B + 4 * S + 76 * E * K
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    achieved = alg / (ms * 1e-3) / 1e9
    return {"value": E * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "tape_bytes_resident": tape.nbytes(), "packed_bytes": int(tape.packed.numel()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_unit": alg / (E * K), "kernel": "adc_replay_packed_kernel"},
            "csr_kernel": {"ms_per_step": ms_csr, "frac": alg / (ms_csr * 1e-3) / 1e9 / peak,
                           "kernel": "adc_replay_implicit_kernel"},
            "note": "tape-driven step (parity mode): pre-drawn volumes / competitor bids / uniforms / revenues "
                    "resident in HBM (each form ~5x the L2), same tape every step; packed = one 16-byte aligned "
                    "record per unit fetched by one bulk copy (TMA) into shared memory; csr_kernel = the same "
                    "tape read through its CSR streams with per-lane loads"}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from adcraft_b200 import _capi
    from adcraft_b200.vector_env import VectorBiddingSimulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    table = workload_table()
    env = VectorBiddingSimulation(
        E_ENVS, num_keywords=K_KW, keywords=table, budget=BUDGET, max_days=MAX_DAYS, device=dev,
        seed=SEED, env_base=rank * E_ENVS, n_lanes=args.n_lanes, obs_dtype=torch.float32,
        updater_mask=[True] * K_KW if DRIFT else None, episode_profit=True)
    env.reset()
    env.budget_alias = bool(args.alias)
    bids = torch.full((E_ENVS, K_KW), BID, dtype=torch.float32, device=dev)
    action = {"keyword_bids": bids}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    from adcraft_b200 import metrics as M
    ideal = M.ideal_profit(env)["ideal"]  # [rows, K] per-step ideal profit (adc_ideal_profit kernel)
    reduce_every = max(1, min(MAX_DAYS, args.steps))
    metric = {"vec": None, "reduces": 0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def episode_reduce():
        # The only collective on this path: per-env AKNCP / NCP over the window (from the kernels'
        # exact per-keyword profit accumulators, experiment_metrics.py:64-83), summed over ranks:
        # 8 doubles through NCCL
        # (rewards and finished episodes are summed per env by the env tail: episode_reward / episode_count)
        vec = M.episode_summary_vector(env, reduce_every, ideal)
        metric["vec"] = M.reduce_metrics(vec)
        metric["reduces"] += 1

    if args.replay_only:
        emit({"replay": run_replay_leg(env, table, dev, torch, steps=max(args.steps, 3))})
        return 0
    if args.explicit:  # experiment: the reference's DEFAULT env (ExplicitKeyword, config 1) vectorised
        from adcraft_b200 import keywords as kwm
        del env
        Kx = args.keywords if args.keywords != 100 else 10
        xtable = kwm.sample_random_keywords(Kx, np.random.default_rng(0))
        xenv = VectorBiddingSimulation(E_ENVS, num_keywords=Kx, keywords=xtable, budget=1000.0, max_days=MAX_DAYS,
                                       device=dev, seed=SEED, env_base=rank * E_ENVS, obs_dtype=torch.float32)
        xenv.reset()
        xb = torch.from_numpy(np.round(np.random.default_rng(1).uniform(0.01, 3.0, (E_ENVS, Kx)), 2).astype(np.float32)).to(dev)
        xact = {"keyword_bids": xb}
        for _ in range(3):
            xo = xenv.step(xact)[0]
        torch.cuda.synchronize(dev)
        n = max(args.steps, 3)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for i in range(n):
            starts[i].record()
            xo = xenv.step(xact)[0]
            stops[i].record()
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops)) / n
        emit({"explicit_default_env": {
            "envs": E_ENVS, "keywords": Kx, "ms_per_step": ms, "units_per_s": E_ENVS * Kx / (ms * 1e-3),
            "mean_volume": float(np.mean(xtable.vol_mean)), "mean_impressions": float(xo["impressions"].float().mean()),
            "note": "ExplicitKeyword set from sample_random_keywords(default_rng(0)), bids U[0.01,3.00], budget 1000 "
                    "(binds for many envs: those take the exact serial kernel)"}})
        return 0
    if args.agents > 1:  # experiment: BASELINE config 4, A bidders inside every auction
        from adcraft_b200.multi_agent import SharedAuctionSimulation
        del env
        A, W = args.agents, E_ENVS
        sim = SharedAuctionSimulation(A, W, num_keywords=K_KW, keywords=table, budget=BUDGET, max_days=MAX_DAYS,
                                      device=dev, seed=SEED, env_base=rank * W, obs_dtype=torch.float32)
        sim.reset()
        sbids = (0.50 + 0.05 * torch.arange(A, device=dev, dtype=torch.float32)).view(1, A, 1).expand(W, A, K_KW).contiguous()
        for _ in range(3):
            sobs = sim.step(sbids)[0]
        torch.cuda.synchronize(dev)
        n = max(args.steps, 3)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for i in range(n):
            starts[i].record()
            sobs = sim.step(sbids)[0]
            stops[i].record()
        torch.cuda.synchronize(dev)
        ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops)) / n
        act = {"keyword_bids": sbids.reshape(W * A, K_KW)}
        for i in range(n):  # A/B: the rival floor as a torch pass + table instead of in the kernel
            starts[i].record()
            sim.vec.step(act, floor_cents=sim.rival_floor_cents(sbids).view(W * A, K_KW))
            stops[i].record()
        torch.cuda.synchronize(dev)
        ms_kernel = sum(a.elapsed_time(b) for a, b in zip(starts, stops)) / n
        winners = (sobs["impressions"] > 0).sum(dim=1)
        emit({"shared_auction": {
            "worlds": W, "agents": A, "keywords": K_KW, "ms_per_step": ms, "ms_per_step_floor_in_torch": ms_kernel,
            "bidder_units_per_s": W * A * K_KW / (ms * 1e-3), "auction_units_per_s": W * K_KW / (ms * 1e-3),
            "max_winners_per_auction_unit": int(winners.max()),
            "note": "one launch over worlds*A bidder rows; every unit finds its highest rival among the A bid rows of "
                    "its world inside the kernel (unit_floor); bids = 0.50 + 0.05*agent"}})
        return 0

    # ---- device-timed steps, inputs resident in HBM --------------------------------------
    sampler = ClockSampler(torch, dev) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        obs, reward, term, trunc, _ = env.step(action)
    for _ in range(2):  # NCCL connects lazily: establish the all-reduce path before the timed region
        episode_reduce()
    metric["reduces"] = 0
    barrier()
    if sampler is not None:
        sampler.mark()
    lib = _capi.load()
    lib.adc_launch_count(1)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()  # evict the 126 MB L2 between steps (outside the event pair)
        starts[i].record()
        obs, reward, term, trunc, _ = env.step(action)
        if (i + 1) % reduce_every == 0:
            episode_reduce()
        stops[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.adc_launch_count(0))
    if sampler is not None and sampler.count_since_mark() < 5:
        # the timed region was shorter than a few NVML polls: keep the GPU under the identical load
        # (same steps, untimed) until the sampler has seen it
        t_up = time.perf_counter()
        while sampler.count_since_mark() < 5 and time.perf_counter() - t_up < 0.25:
            for _ in range(20):
                env.step(action)
            torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    units_total = E_ENVS * K_KW * args.steps * n_gpus
    value = units_total / (dev_ms * 1e-3)
    ms_per_step = dev_ms / args.steps

    # ---- end to end through the public API with HOST buffers ------------------------------
    # step_host_pipelined = adc_step_host: per chunk of envs H2D copy -> kernels -> compact rows ->
    # D2H copy on the chunk's stream; returns when the observations are in host memory.
    bids_host = torch.full((E_ENVS, K_KW), BID, dtype=torch.float32).pin_memory()

    def timed_host(fn):
        for _ in range(max(args.warmup, 10)):  # (the first calls of a path pin its buffers and fault their pages in)
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return units_total / float(te.item())

    e2e_value = timed_host(lambda: env.step_host(bids_host, mode="auto"))
    auto_mode = "pipelined" if world > 2 else "records"
    h2d = E_ENVS * K_KW * 4
    d2h = (E_ENVS * int(_capi.load().adc_host_row_bytes(K_KW, _capi.F32)) if world > 2 else env.host_record_bytes_per_step()[1])
    e2e_rows = timed_host(lambda: env.step_host_rows(bids_host))
    e2e_records = timed_host(lambda: env.step_host_records(bids_host))
    e2e_pipelined = timed_host(lambda: env.step_host_pipelined(bids_host, n_chunks=args.host_chunks))
    e2e_zero_copy = timed_host(lambda: env.step_host(bids_host))
    e2e_staged = timed_host(lambda: env.step_host(bids_host, zero_copy=False))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- replay (tape-driven) leg: the HBM-bound kernel of the path ---------------------------
    replay = None
    if n_gpus == 1 and not args.no_replay:
        replay = run_replay_leg(env, table, dev, torch)

    # ---- roofline of the dominant kernel --------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = E_ENVS * K_KW * B_UNIT
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    cap = ncu_capture_of_hot_kernel()
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": cap["dram_bytes"] if cap else None,
        "kernel": HOT_KERNEL, "peak_source": peak_src, "algorithmic_bytes_per_unit": B_UNIT,
        "issue": {"warp_instructions_per_launch": cap["warp_instructions"] if cap else None,
                  "warp_instructions_per_unit": cap["warp_instructions"] / (E_ENVS * K_KW) if cap else None,
                  "frac_of_issue_peak": (cap["warp_instructions"] / (ms_per_step * 1e-3) / (148 * 4 * 1.965e9)
                                         if cap else None),
                  "note": "instruction count read from the committed ncu capture (profiles/r02_flat2_c2_ncu.json); "
                          "peak = 148 SMs x 4 schedulers x 1965 MHz"},
        "note": "free-running mode decides every auction by a bit-sliced uniform and draws one price per click and "
                "one revenue per conversion from Philox: the kernel is instruction-issue-bound, not HBM-bound "
                "(profiles/r02_summary.md, DESIGN.md); the step = this kernel + a serial-queue kernel that exits "
                "when no budget binds",
    }

    # ---- CPU baseline on this box's host cores (bounded sample) ----------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_block(6, 1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i32+f32", "data": "synthetic",
        "config": config_dict(n_gpus),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": f"VectorBiddingSimulation.step_host(pinned bids, mode='auto') = '{auto_mode}' at {n_gpus} rank(s) per "
                       "host.  records: one launch reads the pinned float32 bids over UVA and writes every unit's observation "
                       "as one aligned 16-byte record (uint16 counts + float32 money) plus the env scalars straight "
                       "into pinned host memory; zero_copy: the same with the five int32 / float32 arrays, 20 B per unit; pipelined "
                       "(adc_step_host): chunks of envs on their own streams, cudaMemcpyAsync in, kernels, compact rows "
                       "(uint16 counts + float32 money, 14 B per unit), cudaMemcpyAsync out.  Either way the call "
                       "returns when the observations are in host memory",
                "other_paths": {f"pipelined_adc_step_host_{args.host_chunks}_chunks": e2e_pipelined,
                                "rows_written_over_uva": e2e_rows, "unit_records_over_uva": e2e_records,
                                "zero_copy_uva_int32_arrays": e2e_zero_copy,
                                "staged_single_copy_int32": e2e_staged},
                "host_ceiling": "tools/host_ceiling.py, 8 ranks of plain cudaMemcpyAsync on one box: 101 GB/s aggregate "
                                "for 8.2 MB int32 blocks (5.05e9 units/s), 138 GB/s for compact rows + bids (7.55e9): the "
                                "shared host memory system, not the GPUs, bounds e2e at 8 GPUs (profiles/r02_host_ceiling_n8.json)"},
        "collective": {"all_reduces_in_timed_region": metric["reduces"], "every_steps": reduce_every,
                       "what": "AKNCP / NCP summary vector (8 doubles) from the kernels' per-keyword profit accumulators",
                       "summary": M.summarize(metric["vec"]) if metric["vec"] is not None else None},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "replay": replay,
        "cpu_binding": numa,
        "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
        "auctions_per_sec": value * MEAN_VOLUME,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------
# C5: large sweep with PPO rollout collection and the NCCL AKNCP / NCP reduce
# --------------------------------------------------------------------------------------------
def run_c5(args):
    """BASELINE.json configs[4]: 10 000 keywords x 131 072 envs per GPU (1M envs on 8), actions from
    a small MLP policy replica on every rank (Gaussian head, PPO-style rollout collection: actions,
    log-probabilities, values and rewards are stored per step; observations are consumed in place
    from the kernel-written flat rows -- 26 GB per step cannot be buffered), and the per-iteration
    AKNCP / NCP all-reduce over NCCL.  Every step of the timed region = policy forward + sampling +
    env step + rollout bookkeeping; the iteration ends with the metric reduce inside the region."""
    import torch
    import torch.distributed as dist
    from adcraft_b200 import _capi, metrics as M
    from adcraft_b200.vector_env import VectorBiddingSimulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    E, K = args.envs, args.keywords
    T = max(args.steps, 1)
    table = workload_table()
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=BUDGET, max_days=MAX_DAYS, device=dev,
                                  seed=SEED, env_base=rank * E, obs_dtype=torch.float32, episode_profit=True,
                                  flat_obs=True)
    env.reset()
    ideal = M.ideal_profit(env)["ideal"]
    torch.manual_seed(1234)  # the same policy replica on every rank
    H = 64
    w1 = (torch.randn(5 * K + 2, H, device=dev) * 0.01)
    w2 = (torch.randn(H, K, device=dev) * 0.01)
    wv = (torch.randn(H, 1, device=dev) * 0.01)
    log_std = torch.full((K,), -1.5, device=dev)
    buf_logp = torch.empty(T, E, device=dev)
    buf_val = torch.empty(T, E, device=dev)
    buf_rew = torch.empty(T, E, device=dev)
    reward_sum = torch.zeros((), dtype=torch.float64, device=dev)
    episodes = torch.zeros((), dtype=torch.float64, device=dev)
    bids = torch.empty(E, K, device=dev)
    noise = torch.empty(E, K, device=dev)
    chunk = 8192  # env rows per policy pass: bounds the temporaries

    def policy_and_step(i):
        flat = env.flat_observation()  # [E, 5K+2], written by the previous step's kernels
        noise.normal_()
        for r0 in range(0, E, chunk):
            x = flat[r0:r0 + chunk]
            h = torch.tanh(torch.log1p(x.clamp_min(0.0)) @ w1)
            mu = 0.75 + h @ w2
            b = bids[r0:r0 + chunk]
            torch.addcmul(mu, noise[r0:r0 + chunk], log_std.exp(), out=b)
            buf_logp[i, r0:r0 + chunk] = (-0.5 * noise[r0:r0 + chunk] ** 2 - log_std).sum(1)
            buf_val[i, r0:r0 + chunk] = (h @ wv).squeeze(1)
            b.clamp_(min=0.01)
        obs, reward, term, trunc, _ = env.step({"keyword_bids": bids})
        buf_rew[i] = reward
        reward_sum.add_(reward.sum())
        episodes.add_(term.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for i in range(max(args.warmup, 1)):
            policy_and_step(0)
        vec = M.reduce_metrics(M.episode_summary_vector(env, max(args.warmup, 1), ideal, reward_sum, episodes))
        reward_sum.zero_(); episodes.zero_()
        barrier()
        lib = _capi.load()
        lib.adc_launch_count(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
        ev0.record()
        for i in range(T):
            policy_and_step(i)
        vec = M.reduce_metrics(M.episode_summary_vector(env, T, ideal, reward_sum, episodes))
        ev1.record()
        barrier()
        # the simulator alone on the same bids (env.step between events)
        for i in range(min(T, 3)):
            sim_ev[i][0].record()
            env.step({"keyword_bids": bids})
            sim_ev[i][1].record()
        barrier()
    launches = int(lib.adc_launch_count(0))
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    sim_ms = float(np.mean([a.elapsed_time(b) for a, b in sim_ev[:min(T, 3)]]))
    if rank == 0:
        units = float(E) * K * T * world
        emit({"metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": T,
              "warmup": max(args.warmup, 1), "ms_per_step": ms / T, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "i32+f32", "data": "synthetic",
              "config": {"workload": "C5: 10k dense implicit keywords x 131072 envs per GPU, MLP policy rollout collection "
                                     "+ NCCL AKNCP/NCP reduce per iteration", "envs_per_gpu": E, "keywords": K,
                         "mean_volume": MEAN_VOLUME, "conversion_rate": CVR, "iteration_steps": T,
                         "parallelism": f"env-sharded x{world}, policy replica per rank"},
              "simulator_only": {"ms_per_step": sim_ms, "units_per_s": float(E) * K * world / (sim_ms * 1e-3)},
              "collective": {"all_reduces_in_timed_region": 1, "summary": M.summarize(vec)},
              "gpu_launches": launches,
              "memory_gb_allocated": torch.cuda.max_memory_allocated(dev) / 1e9,
              "mean_step_reward_per_env": float(buf_rew.mean())})
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    global E_ENVS, BUDGET, K_KW, MEAN_VOLUME, CVR, DRIFT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-lanes", type=int, default=0, dest="n_lanes")
    ap.add_argument("--host-chunks", type=int, default=4, dest="host_chunks",
                    help="env chunks (= streams) of the pipelined host round trip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-replay", action="store_true", help="skip the tape-driven (replay) leg")
    ap.add_argument("--replay-only", action="store_true", help="profiling aid: run only the replay leg")
    ap.add_argument("--budget", type=float, default=None,
                    help="experiment only: per-day budget (default C2's 100000, scaled by keywords / 100 so that "
                         "larger keyword sets stay on the budget-free path like C2 does)")
    ap.add_argument("--keywords", type=int, default=100, help="experiment only (C3: 1000)")
    ap.add_argument("--volume", type=int, default=128, help="experiment only: mean volume (C3: 16 / 64)")
    ap.add_argument("--cvr", type=float, default=0.8, help="experiment only: conversion rate (C3: 0.1)")
    ap.add_argument("--drift", action="store_true", help="experiment only: non-stationary (mask all True)")
    ap.add_argument("--envs", type=int, default=4096, help="experiment only: envs per GPU (default = C2's 4096)")
    ap.add_argument("--alias", action="store_true",
                    help="experiment only: ndarray-budget double charge (bsim:102 + :225), as a Box action space yields")
    ap.add_argument("--explicit", action="store_true",
                    help="experiment only: the reference's default ExplicitKeyword env (config 1) vectorised, K=10")
    ap.add_argument("--agents", type=int, default=1,
                    help="experiment only: bidders per shared auction (BASELINE config 4: --agents 8 --envs 65536)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c3ns", "c4", "c5"],
                    help="BASELINE.json configs: c2 (default, the metric's config), c3 very sparse 1000 x 16384, c3ns "
                         "non-stationary sparse, c4 8 bidders per auction x 65536 worlds, c5 10k x 131072 per GPU with "
                         "policy rollout + metric reduce")
    args = ap.parse_args()
    if args.config == "c3":
        args.keywords, args.envs, args.volume, args.cvr = 1000, 16384, 16, 0.1
    elif args.config == "c3ns":
        args.keywords, args.envs, args.volume, args.cvr, args.drift = 1000, 16384, 64, 0.1, True
    elif args.config == "c4":
        args.agents, args.envs = 8, 65536
    elif args.config == "c5":
        args.keywords = 10000
        if args.envs == 4096:
            args.envs = 131072
    global _OUT
    _OUT = _claim_stdout()
    E_ENVS = args.envs
    BUDGET = args.budget if args.budget is not None else 100000.0 * max(1.0, args.keywords / 100.0)
    K_KW, MEAN_VOLUME, CVR, DRIFT = args.keywords, args.volume, args.cvr, args.drift
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c5":
        return run_c5(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
