"""VectorBiddingSimulation: E copies of the reference ``BiddingSimulation`` advanced per launch.

Host-side mirror of ``adcraft/gymnasium_kw_env.py:22-363`` with a leading env axis.  The
constructor keeps the reference's keyword arguments (``keyword_config, num_keywords, budget,
render_mode, loss_threshold, max_days, updater_params, updater_mask``, env:54-65); ``reset`` /
``step`` keep their signatures and return tuples (env:160-269, :271-346); observations are the
same 7-key dict (``gymnasium_kw_utils.py:45-64``) with shapes ``[E, K]`` / ``[E, 1]``.

All arithmetic of ``step`` runs in the CUDA library behind ``include/adcraft_b200.h``; torch
only owns device memory and streams.  There is no CPU path: constructing the env without a
CUDA device or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _capi
from . import keywords as kwmod
from .spaces import get_action_space, get_observation_space
from .tape import DeviceTape

ArrayLike = Union[np.ndarray, torch.Tensor, Sequence[float], float]

DEFAULT_UPDATER_PARAMS = [["vol", 0.03], ["ctr", 0.03], ["cvr", 0.03]]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _sum_list(xs) -> float:
    """rust.sum_list (src/lib.rs:113-116): sequential f64 sum."""
    acc = 0.0
    for v in xs:
        acc += float(v)
    return acc


def _sum_array(xs) -> float:
    """rust.sum_array (src/lib.rs:107-111): ndarray's 8-way unrolled f64 sum."""
    a = [float(v) for v in xs]
    p = [0.0] * 8
    n8 = len(a) // 8 * 8
    for i in range(0, n8, 8):
        for j in range(8):
            p[j] += a[i + j]
    acc = 0.0
    acc += p[0] + p[4]
    acc += p[1] + p[5]
    acc += p[2] + p[6]
    acc += p[3] + p[7]
    for v in a[n8:]:
        acc += v
    return acc


class VectorBiddingSimulation:
    """E independent bidding environments on one GPU (or one rank's shard of them)."""

    metadata = {"render_modes": ["ansi"]}

    def __init__(
        self,
        num_envs: int,
        keyword_config: Optional[Dict] = None,
        num_keywords: int = 10,
        budget: float = 1000.0,
        render_mode: Optional[str] = None,
        loss_threshold: float = 10000.0,
        max_days: int = 60,
        updater_params: List[List] = DEFAULT_UPDATER_PARAMS,
        updater_mask: Optional[List[bool]] = None,
        *,
        device: Union[str, torch.device] = "cuda",
        seed: Optional[int] = None,
        keywords: Optional[kwmod.KeywordTable] = None,
        shared_keywords: bool = True,
        obs_dtype: torch.dtype = torch.float32,
        env_base: int = 0,
        n_lanes: int = 0,
        budget_alias: bool = False,
        autoreset: bool = True,
        detail_cap: int = 0,
        env_group: int = 0,
        dynamic_work: bool = True,
        spread_outcomes: Optional[bool] = None,
        f32_ties: bool = False,
        episode_profit: bool = False,
        flat_obs: bool = False,
        serial_ws_bytes: int = 4 << 30,
        serial_hint: bool = True,
        **kwargs,
    ) -> None:
        assert render_mode is None or render_mode in self.metadata["render_modes"], (
            f"Specified render_mode of ({render_mode}) is not in the allowed options of (ansi)")
        self._lib = _capi.load()  # raises if the CUDA library is missing
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available() or self._lib.adc_device_count() < 1:
            raise _capi.AdcError("adcraft_b200 needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.keyword_config = keyword_config
        self.num_keywords = int(num_keywords)
        self.budget = float(budget)
        self.render_mode = render_mode
        self.loss_threshold = float(loss_threshold)
        self.max_days = int(max_days)
        self.updater_params = updater_params
        self.updater_mask = None
        self.num_updates = 0
        # Philox key.  Given here or by reset(seed=...); otherwise drawn from np_random (OS entropy) at
        # the first reset, so that unseeded envs do not share one trajectory.
        self.seed = 0 if seed is None else int(seed)
        self._seed_given = seed is not None
        # numpy >= 2 tie rule for float32 bids (adc_step_args.f32_ties): off = the
        # float64 semantics of the build contract (ties always lose)
        self.f32_ties = bool(f32_ties)
        # [E, K] int64 running sum of every step's exact per-keyword profit (adc_step_out.
        # episode_profit_cents): what AKNCP / NCP are made of, accumulated inside the step kernels
        self.episode_profit = bool(episode_profit)
        # [E, 5K+2] flat observation rows in the reference's FlatArrayWrapper layout, written by the
        # kernels (adc_step_out.flat_obs); `flat_observation()` hands the tensor out, no gather pass
        self.want_flat_obs = bool(flat_obs)
        # cap of the exact serial walk's workspace (one slab of ~370 B x K per resident warp)
        self.serial_ws_cap = int(serial_ws_bytes)
        # envs whose budget bound in their previous step skip the budget-free kernel (adc_scratch.
        # serial_hint): same results, a budget-bound env is evaluated once instead of twice
        self.use_serial_hint = bool(serial_hint)
        self.env_base = int(env_base)
        self.n_lanes = int(n_lanes)
        self._auto_lanes = n_lanes == 0
        self.budget_alias = bool(budget_alias)
        self.autoreset = bool(autoreset)
        self.detail_cap = int(detail_cap)  # > 0: record per-click lists on the exact serial path
        # > 1: consecutive groups of env_group envs are the bidders of ONE auction world and share
        # every draw (see adc_step_args.env_group and multi_agent.SharedAuctionSimulation)
        self.env_group = int(env_group)
        # False: no adc_scratch.work_counter, the hot kernel deals its batches statically (A/B and
        # tests; results are identical either way)
        self.dynamic_work = bool(dynamic_work)
        # adc_step_args.spread_outcomes: None = decided from the keyword table (dense, uneven days), True / False for A/B
        self.spread_outcomes = spread_outcomes
        assert self.env_group <= 1 or self.num_envs % self.env_group == 0
        self.shared_keywords = bool(shared_keywords)
        self.obs_dtype = obs_dtype
        assert obs_dtype in (torch.float32, torch.float64)
        self.action_space = get_action_space(self.num_keywords)
        self.observation_space = get_observation_space(self.num_keywords, self.budget)
        self.single_action_space = self.action_space
        self.single_observation_space = self.observation_space
        self.np_random: Optional[np.random.Generator] = None
        self._keywords_given = keywords
        self.keywords: Optional[kwmod.KeywordTable] = None
        self._have_keywords = False
        self._step_count = 0   # Philox counter word: steps since the last seeded reset
        self._calls = 0        # library calls on this scratch: its parity double-buffers the queues
        self._alloc()
        if updater_mask is not None:
            self.set_updater_mask(updater_mask)

    # ------------------------------------------------------------------ allocation
    def _alloc(self) -> None:
        E, K, dev = self.num_envs, self.num_keywords, self.device
        i32, i64, f64 = torch.int32, torch.int64, torch.float64
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)
        # Everything step_host returns lives in ONE contiguous device block (and one pinned host
        # mirror), so the device->host read of a step is a single copy.
        fdt = self.obs_dtype
        layout = [("impressions", i32, (E, K)), ("buyside_clicks", i32, (E, K)),
                  ("sellside_conversions", i32, (E, K)), ("cost", fdt, (E, K)), ("revenue", fdt, (E, K)),
                  ("reward", f64, (E,)), ("cumulative_profit", f64, (E,)), ("days_passed", i32, (E,)),
                  ("terminated", torch.uint8, (E,)), ("truncated", torch.uint8, (E,))]
        self._block_layout, off = [], 0
        for name, dt, shape in layout:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self._block_layout.append((name, dt, shape, off, nbytes))
            off += (nbytes + 15) // 16 * 16
        self._block_bytes = off
        self._block = torch.zeros(off, dtype=torch.uint8, device=dev)
        self._out = self._views(self._block)
        self._out.update(
            cost_cents=z(E, K, dtype=i64), revenue_cents=z(E, K, dtype=i64),
            remaining_budget=z(E, dtype=f64))
        self._state = dict(
            budget=torch.full((E,), self.budget, dtype=f64, device=dev),
            cum_profit=z(E, dtype=f64), day=z(E, dtype=i32))
        self._scratch = dict(
            serial_list=z(E, dtype=i32), serial_count=z(2, dtype=i32), env_profit=z(E, dtype=i64),
            env_cost=z(E, dtype=i64), env_done=z(E, dtype=i32), work_counter=z(2, dtype=i32))
        if self.episode_profit:
            self._out["episode_profit_cents"] = z(E, K, dtype=i64)
            self._out["episode_reward"] = z(E, dtype=f64)
            self._out["episode_count"] = z(E, dtype=i32)
        if self.want_flat_obs:
            self._out["flat_obs"] = z(E, 5 * K + 2, dtype=fdt)
        # one slab per resident warp of the exact serial walk (28 warps per SM), within the cap
        slab = max(int(self._lib.adc_serial_slab_bytes(K)), 1)
        n_slabs = max(1, min(E, 148 * 28, self.serial_ws_cap // slab))
        # the smallest slab keeps 128 clicked slots per keyword; the warps share whatever the workspace has
        # beyond it, up to 512 per keyword (every day of <= 512 auctions fits whatever its clicks are)
        roomy = slab + (K + 31) // 32 * 32 * (512 - 128) * 4
        self._scratch["serial_ws"] = torch.empty(min(n_slabs * roomy, max(self.serial_ws_cap, n_slabs * slab)),
                                                 dtype=torch.uint8, device=dev)
        self._scratch["serial_hint"] = z(E, dtype=torch.uint8)
        self._detail = None
        if self.detail_cap > 0:
            c = self.detail_cap
            self._detail = dict(costs=z(E, K, c, dtype=f64), rev_per_cost=z(E, K, c, dtype=f64),
                                n_recorded=z(E, K, dtype=i32), volume_seen=z(E, K, dtype=f64),
                                lane_clicks=z(E, K, _capi.SUBSTEPS, dtype=i32),
                                lane_convs=z(E, K, _capi.SUBSTEPS, dtype=i32))
        # staging for actions that arrive on the host or in another layout: allocated on first use
        self._bids_dev = {torch.float32: None, torch.float64: None}
        self._budget_dev = {torch.float32: None, torch.float64: None}
        self._mask_dev: Optional[torch.Tensor] = None
        self._host: Dict[str, torch.Tensor] = {}
        self._host_ptrs = None
        self._args = _capi.StepArgs()
        self._args_sig = None
        self._spread_hint = None
        self._obs_cache = None
        self._result_cache = None
        self._host_view_cache = None

    def _call(self, fn, *args) -> None:
        """Run a library entry point with the env's device current: the ABI takes only a stream
        handle and launches on the calling thread's current device (adc_step_args.device makes a
        mismatch an error instead of an illegal address)."""
        if getattr(fn, "__name__", "") != "adc_step_host":
            self._hp_dirty = True  # the pipelined host step must wait for this stream's work once
        if torch.cuda.current_device() == self.device.index:
            _capi.check(fn(*args))
        else:
            with torch.cuda.device(self.device):
                _capi.check(fn(*args))

    def _views(self, block: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {name: block[o:o + n].view(dt).view(shape) for name, dt, shape, o, n in self._block_layout}

    # ------------------------------------------------------------------ keywords / reset
    def set_updater_mask(self, new_updater_mask: List[bool]) -> None:
        """Replace updater mask (env:105-112).  Only the mask changes: parameters drifted so far are
        kept, like the reference, which swaps ``self.updater_mask`` and nothing else."""
        assert len(new_updater_mask) == self.num_keywords, (
            f"Updater mask length ({len(new_updater_mask)})\n"
            + f"must match number of keywords ({self.num_keywords}) to be applied.")
        self.updater_mask = [bool(m) for m in new_updater_mask]
        self.num_updates = int(np.sum(self.updater_mask))
        new = torch.tensor(self.updater_mask, dtype=torch.uint8, device=self.device)
        if self._mask_dev is None:
            self._mask_dev = new
        else:
            self._mask_dev.copy_(new)  # in place: the pointer in the argument block stays valid
        if self._have_keywords and self._kw_stride == 0:
            # drift needs per-env parameter copies: broadcast the CURRENT device values
            E, K = self.num_envs, self.num_keywords
            self._kw_dev = {n: t.reshape(1, K).expand(E, K).contiguous() for n, t in self._kw_dev.items()}
            self._kw_stride = K

    def _install_keywords(self, table: kwmod.KeywordTable) -> None:
        E, K = self.num_envs, self.num_keywords
        assert table.K == K, f"keyword table has K={table.K}, env has {K}"
        per_env = table.per_env or self.updater_mask is not None
        cols = {}
        for name in kwmod.PARAM_NAMES:
            a = np.asarray(getattr(table, name), dtype=np.float64)
            if per_env and a.ndim == 1:
                a = np.broadcast_to(a, (E, K))
            if a.ndim == 2:
                assert a.shape == (E, K), f"{name}: expected {(E, K)}, got {a.shape}"
            cols[name] = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        for name in ("max_bidders", "participation"):  # the class-default ImplicitKeyword's bidder model
            a = getattr(table, name, None)
            if a is not None:
                a = np.asarray(a, dtype=np.float64)
                if per_env and a.ndim == 1:
                    a = np.broadcast_to(a, (E, K))
                cols[name] = torch.from_numpy(np.ascontiguousarray(a).copy()).to(self.device)
        self._kw_dev = cols
        self._kw_stride = K if per_env else 0
        self.keywords = table
        self.kind = table.kind
        self._have_keywords = True

    def install_device_keywords(self, cols: Dict[str, torch.Tensor], kind: int = kwmod.IMPLICIT) -> None:
        """Use per-env keyword parameters that already live on the device ([E, K] float64 tensors,
        e.g. from keywords.sample_implicit_keywords_device) without a host round trip."""
        E, K = self.num_envs, self.num_keywords
        for n in kwmod.PARAM_NAMES:
            t = cols[n]
            assert t.shape == (E, K) and t.dtype == torch.float64 and t.device == self.device, n
        self._kw_dev = {n: cols[n].contiguous() for n in kwmod.PARAM_NAMES}
        self._kw_stride = K
        self.kind = kind
        self.keywords = kwmod.KeywordTable(kind, *[self._kw_dev[n][:1].cpu().numpy() for n in kwmod.PARAM_NAMES])
        self._have_keywords = True

    def bidding_outcomes(self, e: int = 0) -> List[dict]:
        """Per-keyword ``BiddingOutcomes`` dicts (bidding_simulation.py:10-38) of env e for the last
        step, incl. the per-click lists; needs ``detail_cap > 0`` and a step run with
        ``force_serial=True`` and ``n_lanes=1`` (only the exact serial walk records them)."""
        assert self._detail is not None, "construct the env with detail_cap > 0"
        o, d = self._out, self._detail
        K = self.num_keywords
        n = d["n_recorded"][e].cpu().numpy()
        clicks = o["buyside_clicks"][e].cpu().numpy()
        if (clicks > n).any():
            raise _capi.AdcError(
                f"bidding_outcomes: a keyword had {int(clicks.max())} clicks but detail_cap is "
                f"{self.detail_cap}; construct the env with a larger detail_cap")
        costs, rpc = d["costs"][e].cpu().numpy(), d["rev_per_cost"][e].cpu().numpy()
        vol = d["volume_seen"][e].cpu().numpy()
        lane_b, lane_s = d["lane_clicks"][e].cpu().numpy(), d["lane_convs"][e].cpu().numpy()
        imp = o["impressions"][e].cpu().numpy()
        conv = o["sellside_conversions"][e].cpu().numpy()
        bids = self._last_bids[e].double().cpu().numpy()
        out = []
        for k in range(K):
            c, r = costs[k, :n[k]], rpc[k, :n[k]]
            revs = r[r > 0]
            # profit accumulates lane by lane (combine_outcomes, bsim:138-139), each lane's being
            # rust.sum_array(revenues) - rust.sum_list(costs) (bsim:117)
            profit, ib, is_ = 0.0, 0, 0
            for t in range(_capi.SUBSTEPS):
                nb, ns = int(lane_b[k, t]), int(lane_s[k, t])
                profit += _sum_array(revs[is_:is_ + ns]) - _sum_list(c[ib:ib + nb])
                ib, is_ = ib + nb, is_ + ns
            out.append(dict(
                bid=float(np.round(np.maximum(bids[k], 0.01), 2)), impressions=int(imp[k]),
                impression_share=float(imp[k] / vol[k]) if vol[k] > 0 else 0.0,
                buyside_clicks=int(clicks[k]), costs=c.tolist(),
                sellside_conversions=int(conv[k]), revenues=revs.tolist(),
                revenues_per_cost=r.tolist(), profit=float(profit)))
        return out

    def flat_observation(self) -> torch.Tensor:
        """[E, 5K+2] observation rows in the reference's flat layout (wrappers/flat_array.py:44-87),
        written by the step kernels themselves; needs ``flat_obs=True`` at construction.  The env's
        own buffer: overwritten by the next step."""
        assert self.want_flat_obs, "construct the env with flat_obs=True"
        return self._out["flat_obs"]

    def keyword_params(self) -> Dict[str, np.ndarray]:
        """Current (possibly drifted) keyword parameters, host copies."""
        return {n: t.cpu().numpy() for n, t in self._kw_dev.items()}

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Reset every env (env:271-346).  Keywords are re-sampled when a seed is given or none
        exist yet; without a seed the (drifted) keywords persist, as in the reference."""
        if seed is not None or self.np_random is None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        if seed is not None or not self._have_keywords:
            if self._keywords_given is not None:
                table = self._keywords_given
            elif self.keyword_config is not None:
                table = kwmod.sample_implicit_keywords_from_quantiles(
                    self.num_keywords, self.np_random, self.keyword_config,
                    num_envs=None if self.shared_keywords else self.num_envs)
            else:
                table = kwmod.sample_random_keywords(
                    self.num_keywords, self.np_random,
                    num_envs=None if self.shared_keywords else self.num_envs)
            self._install_keywords(table)
        if seed is not None:
            # a seeded reset replays: same key, Philox step counter rewound (the scratch parity,
            # self._calls, keeps running)
            self.seed = int(seed)
            self._seed_given = True
            self._step_count = 0
        elif not self._seed_given:
            self.seed = int(self.np_random.integers(0, 2 ** 63 - 1))
            self._seed_given = True
        if options:
            self.max_days = options.get("max_days", self.max_days)
            rm = options.get("render_mode", self.render_mode)
            if rm is None or rm in self.metadata["render_modes"]:
                self.render_mode = rm
            self.loss_threshold = options.get("loss_threshold", self.loss_threshold)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._call(self._lib.adc_reset_envs,
                   self.num_envs, None, self._state["cum_profit"].data_ptr(), self._state["day"].data_ptr(),
                   C.c_void_p(stream))
        for k in ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue",
                  "cumulative_profit", "days_passed", "reward", "flat_obs", "episode_profit_cents"):
            if k in self._out:
                self._out[k].zero_()
        return self._obs(), {"keyword_params": self.keywords.describe()}

    def reset_envs(self, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Reset the episode state of the envs with ``mask[e] != 0`` only (what a vector front end
        without auto-reset does per finished sub-env); keywords persist like ``reset()`` without a
        seed.  Returns the observation dict, zeroed on the reset rows."""
        m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        assert m.shape == (self.num_envs,)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._call(self._lib.adc_reset_envs,
                   self.num_envs, m.data_ptr(), self._state["cum_profit"].data_ptr(), self._state["day"].data_ptr(),
                   C.c_void_p(stream))
        rows = m.bool()
        for k in ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue",
                  "cumulative_profit", "days_passed", "reward", "flat_obs"):
            if k in self._out:
                self._out[k][rows] = 0
        return self._obs()

    # ------------------------------------------------------------------ step
    def _obs(self) -> Dict[str, torch.Tensor]:
        if self._obs_cache is not None:
            return self._obs_cache
        o = self._out
        self._obs_cache = dict(
            impressions=o["impressions"], buyside_clicks=o["buyside_clicks"], cost=o["cost"],
            sellside_conversions=o["sellside_conversions"], revenue=o["revenue"],
            cumulative_profit=o["cumulative_profit"].view(-1, 1),
            days_passed=o["days_passed"].view(-1, 1))
        return self._obs_cache

    def _stage(self, x: ArrayLike, store: Dict[torch.dtype, torch.Tensor], shape) -> torch.Tensor:
        """Bring an action array onto the device (f32/f64 kept) without a per-step allocation."""
        if isinstance(x, torch.Tensor) and x.device == self.device and x.is_contiguous() \
                and x.dtype in store and tuple(x.shape) == tuple(shape):
            return x
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        if x.dtype not in store:
            x = x.to(torch.float64 if x.dtype == torch.float64 else torch.float32)
        if store[x.dtype] is None:
            store[x.dtype] = torch.zeros(shape, dtype=x.dtype, device=self.device)
        dst = store[x.dtype]
        dst.copy_(x.reshape(shape) if x.numel() == dst.numel() else x.expand(shape), non_blocking=True)
        return dst

    def _fill_args(self, bids: torch.Tensor, budget: Optional[torch.Tensor], force_serial: bool):
        """Per-step fields only; the pointer block is rebuilt when something structural changed."""
        a = self._args
        sig = (self.kind, self._kw_stride, id(self._kw_dev), self.max_days, self.loss_threshold,
               None if self._mask_dev is None else self._mask_dev.data_ptr(), self.num_updates,
               tuple(float(p[1]) for p in self.updater_params), self.autoreset, self.n_lanes,
               self.seed, self.env_base, self.f32_ties)
        if sig != self._args_sig:
            self._fill_static_args()
            self._args_sig = sig
        a.step = self._step_count & 0xFFFFFFFF
        a.parity = self._calls & 0xFFFFFFFF
        a.budget_alias = int(self.budget_alias)
        a.force_serial = int(force_serial)
        a.env_group = self.env_group
        a.spread_outcomes = self._spread_outcomes()
        a.floor_cents = None
        a.bids = bids.data_ptr()
        a.bids_dtype = _capi.F64 if bids.dtype == torch.float64 else _capi.F32
        a.budget_in = _ptr(budget)
        return a

    def _spread_outcomes(self) -> int:
        """adc_step_args.spread_outcomes from the keyword table: worth it for dense, uneven days (a mean volume
        of at least three 32-auction groups and a spread of at least half a group), see the header."""
        if self.spread_outcomes is not None:
            return int(bool(self.spread_outcomes))
        if self._spread_hint is None:
            kw = getattr(self, "_kw_dev", None)
            if self.kind != kwmod.IMPLICIT or not kw:
                self._spread_hint = 0
            else:
                self._spread_hint = int(float(kw["vol_mean"].mean()) >= 96.0 and float(kw["vol_std"].max()) >= 16.0)
        return self._spread_hint

    def _fill_static_args(self) -> None:
        self._spread_hint = None  # (the keyword table may be another one)
        a, o, s, st = self._args, self._out, self._scratch, self._state
        E, K = self.num_envs, self.num_keywords
        a.E, a.env_base, a.seed = E, self.env_base, self.seed & 0xFFFFFFFFFFFFFFFF
        a.device = self.device.index
        a.f32_ties = int(self.f32_ties)
        a.n_lanes = self.n_lanes
        a.autoreset = int(self.autoreset)
        kw = a.kw
        kw.kind, kw.K, kw.env_stride = self.kind, K, self._kw_stride
        for n in kwmod.PARAM_NAMES:
            setattr(kw, n, self._kw_dev[n].data_ptr())
        kw.max_bidders = _ptr(self._kw_dev.get("max_bidders"))
        kw.participation = _ptr(self._kw_dev.get("participation"))
        kw.impression_thresh = self.keywords.impression_thresh
        a.env.budget, a.env.cum_profit, a.env.day = (st["budget"].data_ptr(), st["cum_profit"].data_ptr(),
                                                     st["day"].data_ptr())
        a.env.max_days, a.env.loss_threshold = self.max_days, self.loss_threshold
        if self.updater_mask is not None:
            a.drift.mask, a.drift.num_updates = self._mask_dev.data_ptr(), self.num_updates
            a.drift.mag = (C.c_double * 3)(*[float(p[1]) for p in self.updater_params])
        else:
            a.drift.mask, a.drift.num_updates = None, 0
        out = a.out
        out.impressions, out.clicks = o["impressions"].data_ptr(), o["buyside_clicks"].data_ptr()
        out.conversions = o["sellside_conversions"].data_ptr()
        out.cost, out.revenue = o["cost"].data_ptr(), o["revenue"].data_ptr()
        out.float_dtype = _capi.F64 if self.obs_dtype == torch.float64 else _capi.F32
        out.cost_cents, out.revenue_cents = o["cost_cents"].data_ptr(), o["revenue_cents"].data_ptr()
        out.reward, out.obs_cum_profit = o["reward"].data_ptr(), o["cumulative_profit"].data_ptr()
        out.obs_days = o["days_passed"].data_ptr()
        out.terminated, out.truncated = o["terminated"].data_ptr(), o["truncated"].data_ptr()
        out.remaining_budget = o["remaining_budget"].data_ptr()
        sc = a.scratch
        if self.kind != kwmod.IMPLICIT and "unit_cost_f64" not in s:  # un-rounded cost sums: explicit / multi-bidder
            s["unit_cost_f64"] = torch.zeros(E, K, dtype=torch.float64, device=self.device)
        for n in ("serial_list", "serial_count", "env_profit", "env_cost", "env_done"):
            setattr(sc, n, s[n].data_ptr())
        sc.unit_cost_f64 = _ptr(s.get("unit_cost_f64"))
        sc.work_counter = s["work_counter"].data_ptr() if self.dynamic_work else None
        sc.serial_ws, sc.serial_ws_bytes = s["serial_ws"].data_ptr(), s["serial_ws"].numel()
        sc.serial_hint = s["serial_hint"].data_ptr() if self.use_serial_hint else None
        if self.env_group > 1 and self.kind == kwmod.IMPLICIT and "outbid_mask" not in s:
            s["outbid_mask"] = torch.zeros(E, K, dtype=torch.uint8, device=self.device)  # shared auctions: units finished by the pre-pass
        sc.outbid_mask = _ptr(s.get("outbid_mask"))
        out.episode_profit_cents = _ptr(o.get("episode_profit_cents"))
        out.episode_reward = _ptr(o.get("episode_reward"))
        out.episode_count = _ptr(o.get("episode_count"))
        out.rows = None
        out.flat_obs = _ptr(o.get("flat_obs"))
        if self._detail is not None:
            a.detail.cap = self.detail_cap
            for n in ("costs", "rev_per_cost", "n_recorded", "volume_seen", "lane_clicks", "lane_convs"):
                setattr(a.detail, n, self._detail[n].data_ptr())

    def _prepare(self, action: Dict[str, ArrayLike]):
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        E, K = self.num_envs, self.num_keywords
        bids = self._stage(action["keyword_bids"], self._bids_dev, (E, K))
        budget = action.get("budget") if isinstance(action, dict) else None
        if budget is not None:
            budget = self._stage(budget, self._budget_dev, (E,))
            if budget.dtype != bids.dtype:  # one dtype tag covers both in the ABI
                budget = self._stage(budget.to(bids.dtype), self._budget_dev, (E,))
        return bids, budget

    def step(self, action: Dict[str, ArrayLike], *, force_serial: bool = False,
             floor_cents: Optional[torch.Tensor] = None):
        """One env step for all E envs (env:160-269).  Returns device tensors.  ``floor_cents``
        ([E, K] int32 on the device): highest rival bid per unit for shared auctions.

        The returned observation / reward / flag tensors are the env's own output buffers: every
        call overwrites them in place (no per-step allocation).  A caller that keeps a step's
        results past the next ``step`` (a rollout buffer) must ``clone()`` them or copy them into its
        own storage; the vector adapters' ``copy=True`` does that."""
        bids, budget = self._prepare(action)
        self._last_bids = bids
        a = self._fill_args(bids, budget, force_serial)
        if floor_cents is not None:
            assert floor_cents.dtype == torch.int32 and floor_cents.is_contiguous() and floor_cents.device == self.device
            assert tuple(floor_cents.shape) == (self.num_envs, self.num_keywords)
            a.floor_cents = floor_cents.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._call(self._lib.adc_step_philox, C.byref(a), C.c_void_p(stream))
        self._step_count += 1
        self._calls += 1
        self._hp_dirty = True
        return self._result()

    def step_replay(self, action: Dict[str, ArrayLike], tape: DeviceTape, *, force_serial: bool = False):
        """Parity mode: the same step fed by pre-drawn volumes / bids / uniforms / revenues."""
        bids, budget = self._prepare(action)
        self._last_bids = bids
        a = self._fill_args(bids, budget, force_serial)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        t = tape.c_struct()
        self._call(self._lib.adc_step_replay, C.byref(a), C.byref(t), C.c_void_p(stream))
        self._step_count += 1
        self._calls += 1
        return self._result()

    def _result(self):
        if self._result_cache is None:
            o = self._out
            self._result_cache = (self._obs(), o["reward"], o["terminated"].view(torch.bool),
                                  o["truncated"].view(torch.bool))
        return (*self._result_cache, {"step": self._step_count})

    # ------------------------------------------------------------------ host round trip (e2e)
    # ------------------------------------------------------------------ pipelined host round trip
    def _host_pipeline(self, n_chunks: int):
        """Everything adc_step_host needs, built once per (argument block, chunk count): the chunk
        argument blocks (every pointer offset to the chunk's first env, own scratch counters), one
        stream per chunk, the device and pinned-host row blocks."""
        E, K, dev = self.num_envs, self.num_keywords, self.device
        A = max(self.env_group, 1)
        n_chunks = max(1, min(int(n_chunks), E // A))
        key = (self._args_sig, n_chunks)
        hp = getattr(self, "_hp", None)
        if hp is not None and hp["key"] == key:
            return hp
        fdt = _capi.F64 if self.obs_dtype == torch.float64 else _capi.F32
        row_bytes = int(self._lib.adc_host_row_bytes(K, fdt))
        if hp is None or hp["rows_host"].shape != (E, row_bytes):
            rows_dev = torch.zeros(E, row_bytes, dtype=torch.uint8, device=dev)
            rows_host = torch.zeros(E, row_bytes, dtype=torch.uint8).pin_memory()
            bids_dev = torch.zeros(E, K, dtype=torch.float32, device=dev)
        else:
            rows_dev, rows_host, bids_dev = hp["rows_dev"], hp["rows_host"], hp["bids_dev"]
        counters = torch.zeros(n_chunks, 2, 2, dtype=torch.int32, device=dev)  # [chunk][serial|work][parity]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_chunks)]
        chunks = (_capi.HostChunk * n_chunks)()
        # split the envs into n_chunks runs of whole bidder groups
        groups = E // A
        bounds = [(groups * i // n_chunks) * A for i in range(n_chunks + 1)]
        base = self._args
        ws, ws_bytes = base.scratch.serial_ws, base.scratch.serial_ws_bytes
        slab = max(int(self._lib.adc_serial_slab_bytes(K)), 1)
        per_chunk_ws = (ws_bytes // n_chunks) // slab * slab if ws else 0

        def off(ptr, nbytes):
            return None if not ptr else ptr + nbytes

        for i in range(n_chunks):
            e0, e1 = bounds[i], bounds[i + 1]
            c = chunks[i]
            C.memmove(C.byref(c.args), C.byref(base), C.sizeof(_capi.StepArgs))
            a = c.args
            a.E, a.env_base = e1 - e0, self.env_base + (e0 // A if A > 1 else e0)
            if self._kw_stride:
                for n in kwmod.PARAM_NAMES + ("max_bidders", "participation"):
                    setattr(a.kw, n, off(getattr(base.kw, n), e0 * K * 8))
            a.env.budget, a.env.cum_profit = off(base.env.budget, e0 * 8), off(base.env.cum_profit, e0 * 8)
            a.env.day = off(base.env.day, e0 * 4)
            a.bids, a.bids_dtype = bids_dev.data_ptr() + e0 * K * 4, _capi.F32
            a.budget_in, a.floor_cents = None, None
            o, bo = a.out, base.out
            fb = 8 if self.obs_dtype == torch.float64 else 4
            for n, sz in (("impressions", 4), ("clicks", 4), ("conversions", 4), ("cost", fb), ("revenue", fb),
                          ("cost_cents", 8), ("revenue_cents", 8), ("episode_profit_cents", 8)):
                setattr(o, n, off(getattr(bo, n), e0 * K * sz))
            for n, sz in (("reward", 8), ("obs_cum_profit", 8), ("obs_days", 4), ("terminated", 1), ("truncated", 1),
                          ("remaining_budget", 8), ("episode_reward", 8), ("episode_count", 4)):
                setattr(o, n, off(getattr(bo, n), e0 * sz))
            o.rows = None
            o.flat_obs = off(bo.flat_obs, e0 * (5 * K + 2) * fb)
            sc, bs = a.scratch, base.scratch
            for n, sz in (("serial_list", 4), ("env_profit", 8), ("env_cost", 8), ("env_done", 4), ("serial_hint", 1)):
                setattr(sc, n, off(getattr(bs, n), e0 * sz))
            sc.unit_cost_f64 = off(bs.unit_cost_f64, e0 * K * 8)
            sc.outbid_mask = off(bs.outbid_mask, e0 * K)
            sc.serial_count = counters[i, 0].data_ptr()
            sc.work_counter = counters[i, 1].data_ptr() if self.dynamic_work else None
            sc.acc_impressions = sc.acc_clicks = sc.acc_conversions = None
            sc.serial_ws = off(ws, i * per_chunk_ws) if per_chunk_ws else None
            sc.serial_ws_bytes = per_chunk_ws
            a.detail.costs = a.detail.rev_per_cost = a.detail.n_recorded = a.detail.volume_seen = None
            a.detail.lane_clicks = a.detail.lane_convs = None
            c.rows_dev = rows_dev.data_ptr() + e0 * row_bytes
            c.rows_host = rows_host.data_ptr() + e0 * row_bytes
            c.stream = streams[i].cuda_stream
        views = self._row_views(rows_host)
        self._hp = dict(key=key, chunks=chunks, n=n_chunks, bounds=bounds, streams=streams, counters=counters,
                        rows_dev=rows_dev, rows_host=rows_host, bids_dev=bids_dev, views=views, row_bytes=row_bytes)
        return self._hp

    def _row_views(self, r: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Strided views of a [E, row_bytes] uint8 row block (layout: adc_host_chunk in the header)."""
        K = self.num_keywords
        fbytes = 8 if self.obs_dtype == torch.float64 else 4
        L6 = (6 * K + 7) // 8 * 8
        tail = (L6 + 2 * fbytes * K + 7) // 8 * 8
        u16, fdt_t = torch.uint16, self.obs_dtype
        return dict(
            impressions=r[:, 0:2 * K].view(u16), buyside_clicks=r[:, 2 * K:4 * K].view(u16),
            sellside_conversions=r[:, 4 * K:6 * K].view(u16),
            cost=r[:, L6:L6 + fbytes * K].view(fdt_t), revenue=r[:, L6 + fbytes * K:L6 + 2 * fbytes * K].view(fdt_t),
            reward=r[:, tail:tail + 8].view(torch.float64).view(-1),
            cumulative_profit=r[:, tail + 8:tail + 16].view(torch.float64),
            days_passed=r[:, tail + 16:tail + 20].view(torch.int32),
            terminated=r[:, tail + 20], truncated=r[:, tail + 21], count_overflow=r[:, tail + 22])

    def step_host_rows(self, bids_host: torch.Tensor):
        """``step`` for a CPU-side caller, fused: ONE launch reads the pinned HOST bids (UVA) and the
        warp that finalises an env packs its observation into a compact row (uint16 counts, float
        money: 14 bytes per unit, layout of ``adc_host_chunk``) straight into pinned HOST memory
        (``adc_step_out.rows``), so the PCIe traffic of both directions overlaps the auction work and no
        copy is enqueued.  Returns strided views into the pinned row block, valid until the next
        call; ``count_overflow[e]`` flags an env with a count above 65535."""
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        E, K = self.num_envs, self.num_keywords
        assert (isinstance(bids_host, torch.Tensor) and bids_host.is_pinned() and bids_host.is_contiguous()
                and bids_host.dtype in (torch.float32, torch.float64) and tuple(bids_host.shape) == (E, K)), \
            "step_host_rows takes a pinned, contiguous float [E, K] tensor"
        if getattr(self, "_rows_host", None) is None:
            fdt = _capi.F64 if self.obs_dtype == torch.float64 else _capi.F32
            self._rows_host = torch.zeros(E, int(self._lib.adc_host_row_bytes(K, fdt)), dtype=torch.uint8).pin_memory()
            self._rows_views = self._row_views(self._rows_host)
        a = self._fill_args(bids_host, None, False)
        a.out.rows = self._rows_host.data_ptr()
        stream = torch.cuda.current_stream(self.device)
        try:
            self._call(self._lib.adc_step_philox, C.byref(a), C.c_void_p(stream.cuda_stream))
        finally:
            a.out.rows = None
        self._step_count += 1
        self._calls += 1
        stream.synchronize()
        return self._rows_views

    def step_host_records(self, bids_host: torch.Tensor, budget_host: Optional[torch.Tensor] = None):
        """``step`` for a CPU-side caller, fused, in its leanest form: ONE launch reads the pinned HOST
        bids (UVA) and every unit's observation goes to pinned HOST memory as one aligned 16-byte record
        (``adc_step_out.unit_records``: uint16 impressions / clicks / conversions / flags, float32
        cost / revenue) -- 512 contiguous bytes per warp, full-size PCIe write transactions, 16 bytes per
        unit instead of the 20 of ``mode="zero_copy"``'s five 4-byte streams; the env scalars land in
        pinned host arrays as well.  No copy is enqueued; the int32 / float observation tensors of
        ``step`` stay on the device and are updated too.  Returns strided views into the pinned record
        block, valid until the next call; ``count_overflow[e, k]`` flags a count above 65535."""
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        E, K = self.num_envs, self.num_keywords
        assert self.obs_dtype == torch.float32, "unit records carry float32 money: obs_dtype must be float32"
        assert (isinstance(bids_host, torch.Tensor) and bids_host.is_pinned() and bids_host.is_contiguous()
                and bids_host.dtype in (torch.float32, torch.float64) and tuple(bids_host.shape) == (E, K)), \
            "step_host_records takes a pinned, contiguous float [E, K] tensor"
        if getattr(self, "_rec_host", None) is None:
            rec = torch.zeros(E, K, 16, dtype=torch.uint8).pin_memory()
            # env scalars, one pinned array each: reward f64 | cum_profit f64 | days i32 | terminated u8 | truncated u8
            envT = torch.zeros(24 * E, dtype=torch.uint8).pin_memory()
            u16, f32 = rec.view(torch.uint16), rec.view(torch.float32)
            scal = dict(reward=envT[0:8 * E].view(torch.float64), cumulative_profit=envT[8 * E:16 * E].view(torch.float64),
                        days_passed=envT[16 * E:20 * E].view(torch.int32), terminated=envT[20 * E:21 * E],
                        truncated=envT[21 * E:22 * E])
            self._rec_host, self._rec_env = rec, envT
            self._rec_ptrs = tuple(scal[k].data_ptr() for k in
                                   ("reward", "cumulative_profit", "days_passed", "terminated", "truncated"))
            self._rec_views = dict(
                impressions=u16[:, :, 0], buyside_clicks=u16[:, :, 1], sellside_conversions=u16[:, :, 2],
                count_overflow=u16[:, :, 3], cost=f32[:, :, 2], revenue=f32[:, :, 3],
                reward=scal["reward"], cumulative_profit=scal["cumulative_profit"].view(-1, 1),
                days_passed=scal["days_passed"].view(-1, 1), terminated=scal["terminated"], truncated=scal["truncated"])
        budget = None
        if budget_host is not None:
            budget = self._stage(budget_host.to(bids_host.dtype), self._budget_dev, (E,))
        a = self._fill_args(bids_host, budget, False)
        out = a.out
        saved = (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated)
        (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = self._rec_ptrs
        out.unit_records = self._rec_host.data_ptr()
        stream = torch.cuda.current_stream(self.device)
        try:
            self._call(self._lib.adc_step_philox, C.byref(a), C.c_void_p(stream.cuda_stream))
        finally:
            (out.reward, out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = saved
            out.unit_records = None
        self._step_count += 1
        self._calls += 1
        stream.synchronize()
        return self._rec_views

    def step_host_pipelined(self, bids_host: torch.Tensor, n_chunks: int = 4):
        """``step`` for a CPU-side caller: pinned float32 HOST bids in, HOST observations out, through
        ``adc_step_host``: the envs are cut into ``n_chunks`` runs and every run's host->device copy,
        kernels, row packing and device->host copy go to its own stream, so the copy engines move
        one run while the SMs work on the next (one cudaMemcpyAsync per run and direction).
        Observations come back as compact rows (uint16 counts, float money; adc_host_chunk in the
        header): the returned dict holds strided views into the pinned row block, valid until the
        next call.  ``count_overflow[e]`` flags an env with a count above 65535 (read its exact
        int32 counts from the device arrays then).  The device-side observation tensors of ``step``
        are updated as well."""
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        E, K = self.num_envs, self.num_keywords
        assert (isinstance(bids_host, torch.Tensor) and bids_host.is_pinned() and bids_host.dtype == torch.float32
                and bids_host.is_contiguous() and tuple(bids_host.shape) == (E, K)), \
            "step_host_pipelined takes a pinned, contiguous float32 [E, K] tensor"
        self._fill_args(bids_host, None, False)  # per-step fields + (re)built pointer block
        hp = self._host_pipeline(n_chunks)
        if getattr(self, "_hp_dirty", True):
            torch.cuda.synchronize(self.device)  # work queued on other streams touches the same state
            self._hp_dirty = False
        base, chunks = self._args, hp["chunks"]
        for i in range(hp["n"]):
            a = chunks[i].args
            a.step, a.parity = base.step, base.parity
            a.budget_alias, a.force_serial, a.f32_ties = base.budget_alias, 0, base.f32_ties
            chunks[i].bids_host = bids_host.data_ptr() + hp["bounds"][i] * K * 4
        self._call(self._lib.adc_step_host, chunks, hp["n"])
        self._step_count += 1
        self._calls += 1
        return hp["views"]

    def step_host(self, bids_host: torch.Tensor, budget_host: Optional[torch.Tensor] = None,
                  zero_copy: bool = True, mode: Optional[str] = None):
        """``step`` with HOST buffers: pinned host bids in, pinned host observations out.

        ``mode`` picks how the two transfers are made (default ``None``: ``zero_copy`` decides between
        the first and the last, as in round 1):

        * ``"zero_copy"``  the kernels read the bids from the pinned host buffer and write every
          observation array (int32 counts, float money) into a pinned host block (UVA), 20 bytes per
          unit in five 4-byte streams (B200: 1.65e9 units/s on C2);
        * ``"pipelined"``  ``step_host_pipelined``: chunks of envs on their own streams, copy engines,
          compact rows (uint16 counts): a third fewer bytes and far fewer transactions through the
          host memory system -- the fastest form when SEVERAL ranks share one host (8 x B200: 8.0e9
          units/s, 96 % of what 8 concurrent cudaMemcpyAsync streams of the same rows reach);
        * ``"rows"``       ``step_host_rows``: compact rows written by the kernels over UVA;
        * ``"records"``    ``step_host_records``: one aligned 16-byte record per unit written by the
          kernels over UVA (float32 observations only): the fastest form for ONE process per host
          (B200: 1.91e9 units/s on C2, 88 % of the device-timed rate);
        * ``"auto"``       "pipelined" when torch.distributed runs more than two ranks, else "records"
          ("zero_copy" for float64 observations);
        * ``"staged"``     one H2D copy, the step, one D2H copy of the contiguous int32 block.

        Returns a dict of pinned host tensors (views; valid until the next call).  The compact forms
        return uint16 counts and a ``count_overflow`` flag per env."""
        if mode == "auto":
            import torch.distributed as dist
            # measured on C2 (units/s end to end, records vs pipelined): 1 rank 1.92e9 vs 1.32e9, 2 ranks 3.84e9 vs
            # 2.61e9, 4 ranks 4.57e9 vs 4.68e9, 8 ranks: the host memory system binds and the pipeline's fewer
            # bytes win (8.0e9 vs 6.1e9 for rows written over UVA)
            shared = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 2
            mode = "pipelined" if shared else ("records" if self.obs_dtype == torch.float32 else "zero_copy")
        if mode == "records":
            return self.step_host_records(bids_host, budget_host)
        if mode in ("pipelined", "rows") and budget_host is None and self.kind == kwmod.IMPLICIT:
            return self.step_host_pipelined(bids_host) if mode == "pipelined" else self.step_host_rows(bids_host)
        if mode is not None:
            zero_copy = mode not in ("staged",)
        if not self._host:
            self._host_block = torch.zeros(self._block_bytes, dtype=torch.uint8).pin_memory()
            self._host = self._views(self._host_block)
            self._host_ptrs = None
        if not zero_copy or not (isinstance(bids_host, torch.Tensor) and bids_host.is_pinned()
                                 and bids_host.dtype in (torch.float32, torch.float64)
                                 and bids_host.is_contiguous()
                                 and tuple(bids_host.shape) == (self.num_envs, self.num_keywords)):
            action = {"keyword_bids": bids_host}
            if budget_host is not None:
                action["budget"] = budget_host
            self.step(action)
            self._host_block.copy_(self._block, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return self._host_views()
        assert self._have_keywords, "reset required, need to generate keywords to bid on"
        budget = None
        if budget_host is not None:
            budget = self._stage(budget_host.to(bids_host.dtype), self._budget_dev, (self.num_envs,))
        a = self._fill_args(bids_host, budget, False)
        out, h = a.out, self._host
        saved = (out.impressions, out.clicks, out.conversions, out.cost, out.revenue, out.reward,
                 out.obs_cum_profit, out.obs_days, out.terminated, out.truncated)
        if self._host_ptrs is None:  # pointer values are fixed once the pinned block exists
            self._host_ptrs = tuple(h[k].data_ptr() for k in (
                "impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue", "reward",
                "cumulative_profit", "days_passed", "terminated", "truncated"))
        (out.impressions, out.clicks, out.conversions, out.cost, out.revenue, out.reward,
         out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = self._host_ptrs
        # the exact serial walk keeps its running counts in the (otherwise idle) device observation
        # arrays instead of read-modify-writing host memory across PCIe
        sc = a.scratch
        sc.acc_impressions, sc.acc_clicks, sc.acc_conversions = saved[0], saved[1], saved[2]
        stream = torch.cuda.current_stream(self.device)
        try:
            self._call(self._lib.adc_step_philox, C.byref(a), C.c_void_p(stream.cuda_stream))
        finally:
            (out.impressions, out.clicks, out.conversions, out.cost, out.revenue, out.reward,
             out.obs_cum_profit, out.obs_days, out.terminated, out.truncated) = saved
            sc.acc_impressions = sc.acc_clicks = sc.acc_conversions = None
        self._step_count += 1
        self._calls += 1
        stream.synchronize()
        return self._host_views()

    def _host_views(self) -> Dict[str, torch.Tensor]:
        if self._host_view_cache is None:
            h = dict(self._host)
            h["cumulative_profit"] = h["cumulative_profit"].view(-1, 1)
            h["days_passed"] = h["days_passed"].view(-1, 1)
            self._host_view_cache = h
        return self._host_view_cache

    def host_bytes_per_step(self):
        """(h2d, d2h) bytes moved by step_host for this shape."""
        E, K = self.num_envs, self.num_keywords
        fb = 8 if self.obs_dtype == torch.float64 else 4
        return E * K * 4, self._block_bytes

    def host_record_bytes_per_step(self):
        """(h2d, d2h) bytes moved by step_host(mode="records") for this shape."""
        E, K = self.num_envs, self.num_keywords
        return E * K * 4, E * K * 16 + E * 22

    # ------------------------------------------------------------------ misc reference API
    def render(self) -> Optional[str]:
        if self.render_mode == "ansi":
            r = self._out["reward"]
            return (f"Time step: {int(self._out['days_passed'].max())}/{self.max_days},   "
                    f"Mean profit per env in step: {float(r.mean()):.2f}\n")
        return None

    def close(self) -> None:
        pass

    @property
    def launches(self) -> int:
        return int(self._lib.adc_launch_count(0))
