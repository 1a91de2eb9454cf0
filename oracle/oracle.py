"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (adcraft_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``adcraft_b200`` never does (tests/test_layout.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
SUBSTEPS = 24
IMPLICIT, EXPLICIT, IMPLICIT_MULTI = 0, 1, 2


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    so = os.path.join(_BUILD, "liboracle.so")
    src = os.path.join(_HERE, "adcraft_oracle.c")
    hdr = os.path.join(_HERE, "adcraft_oracle.h")
    tab = os.path.join(_HERE, "neglog_table.inc")
    if (not force and os.path.exists(so)
            and os.path.getmtime(so) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(tab))):
        return
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    name = "liboracle.so" if _cpu_has_fma() else "liboracle_generic.so"
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    L.orc_neglog_u31.restype = C.c_float
    L.orc_neglog_u31.argtypes = [C.c_uint32]
    L.orc_znorm.restype = C.c_float
    L.orc_znorm.argtypes = [C.c_uint32]
    L.orc_exp.restype = C.c_double
    L.orc_exp.argtypes = [C.c_double]
    L.orc_laplace_cents.restype = C.c_int32
    L.orc_laplace_cents.argtypes = [C.c_uint32, C.c_float, C.c_float]
    L.orc_revenue_cents.restype = C.c_int32
    L.orc_revenue_cents.argtypes = [C.c_uint32, C.c_float, C.c_float]
    L.orc_volume.restype = C.c_int64
    L.orc_volume.argtypes = [C.c_uint32, C.c_double, C.c_double]
    L.orc_threshold_sigmoid.restype = C.c_double
    L.orc_threshold_sigmoid.argtypes = [C.c_double] * 4
    L.orc_prob_threshold.restype = C.c_uint32
    L.orc_prob_threshold.argtypes = [C.c_double]
    L.orc_explicit_cost.restype = C.c_double
    L.orc_explicit_cost.argtypes = [C.c_uint32, C.c_double]
    L.orc_sum_array.restype = C.c_double
    L.orc_sum_array.argtypes = [C.c_void_p, C.c_int64]
    L.orc_bid_to_cents.restype = C.c_int32
    L.orc_bid_to_cents.argtypes = [C.c_double]
    L.orc_step_replay.restype = C.c_int
    L.orc_step_philox.restype = C.c_int
    L.orc_step_philox_shared.restype = C.c_int
    L.orc_batch_step.restype = C.c_int
    _lib = L
    return L


# ----------------------------------------------------------------------------- #
# ctypes mirrors of the structs in adcraft_oracle.h
# ----------------------------------------------------------------------------- #
class _Keywords(C.Structure):
    _fields_ = [("kind", C.c_int32), ("K", C.c_int32)] + [
        (n, C.c_void_p) for n in ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std",
                                  "max_bidders", "participation")
    ] + [("impression_thresh", C.c_double)]


class _Tape(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "volume", "comp_off", "comp_cents", "click_off", "u_click", "conv_off", "u_conv",
        "rev_off", "rev_cents", "impr", "cost_off", "cost", "comp_f64")]


class _Record(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("cap_per_kw", C.c_int64)] + [
        (n, C.c_void_p) for n in ("comp_cents", "u_click", "u_conv", "rev_cents", "impr", "cost", "comp_f64",
                                  "n_comp", "n_click", "n_conv", "n_rev", "n_cost")]


class _Result(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "impressions", "clicks", "conversions", "cost", "revenue", "profit",
        "cost_cents", "revenue_cents", "lane_I", "lane_B", "lane_S")] + [
        ("reward", C.c_double), ("remaining_budget", C.c_double), ("lanes_run", C.c_int32)]


class _Batch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("E", C.c_int32), ("K", C.c_int32),
                ("param_env_stride", C.c_int64)] + [
        (n, C.c_void_p) for n in ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std",
                                  "max_bidders", "participation")
    ] + [("impression_thresh", C.c_double), ("drift_mask", C.c_void_p),
         ("drift_mag", C.c_double * 3), ("budget", C.c_void_p), ("cum_profit", C.c_void_p),
         ("day", C.c_void_p), ("max_days", C.c_int32), ("loss_threshold", C.c_double),
         ("seed", C.c_uint64), ("env_base", C.c_uint32), ("step", C.c_uint32),
         ("budget_alias", C.c_int32)]


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


PARAM_NAMES = ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std")


@dataclass
class KeywordSet:
    """SoA keyword parameters for one env: 8 float64 columns of length K."""
    kind: int
    vol_mean: np.ndarray
    vol_std: np.ndarray
    p1: np.ndarray
    p2: np.ndarray
    ctr: np.ndarray
    cvr: np.ndarray
    rev_mean: np.ndarray
    rev_std: np.ndarray
    impression_thresh: float = 0.05
    max_bidders: Optional[np.ndarray] = None     # IMPLICIT_MULTI only (classes:659-662)
    participation: Optional[np.ndarray] = None   # IMPLICIT_MULTI only (classes:663)

    def __post_init__(self):
        for n in PARAM_NAMES:
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float64))
        for n in ("max_bidders", "participation"):
            if getattr(self, n) is not None:
                setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float64))

    @property
    def K(self) -> int:
        return len(self.vol_mean)

    def c_struct(self) -> _Keywords:
        s = _Keywords()
        s.kind, s.K = self.kind, self.K
        for n in PARAM_NAMES:
            setattr(s, n, getattr(self, n).ctypes.data)
        s.max_bidders, s.participation = _p(self.max_bidders), _p(self.participation)
        s.impression_thresh = self.impression_thresh
        return s

    def copy(self) -> "KeywordSet":
        return KeywordSet(self.kind, *[getattr(self, n).copy() for n in PARAM_NAMES],
                          impression_thresh=self.impression_thresh,
                          max_bidders=None if self.max_bidders is None else self.max_bidders.copy(),
                          participation=None if self.participation is None else self.participation.copy())


@dataclass
class Tape:
    """Replay tape of ONE env step (consumption order, CSR per keyword)."""
    volume: np.ndarray
    comp_off: np.ndarray
    comp_cents: np.ndarray
    click_off: np.ndarray
    u_click: np.ndarray
    conv_off: np.ndarray
    u_conv: np.ndarray
    rev_off: np.ndarray
    rev_cents: np.ndarray
    impr: Optional[np.ndarray] = None       # explicit: [K,24]
    cost_off: Optional[np.ndarray] = None
    cost: Optional[np.ndarray] = None
    drift: Optional[np.ndarray] = None      # [3,K] coefficients (vol, ctr, cvr)
    comp_f64: Optional[np.ndarray] = None   # multi-bidder keywords: highest bid per auction (shares comp_off)

    def normalise(self) -> "Tape":
        self.volume = np.ascontiguousarray(self.volume, np.int32)
        for n in ("comp_off", "click_off", "conv_off", "rev_off", "cost_off"):
            v = getattr(self, n)
            if v is not None:
                setattr(self, n, np.ascontiguousarray(v, np.int64))
        for n in ("comp_cents", "rev_cents", "impr"):
            v = getattr(self, n)
            if v is not None:
                setattr(self, n, np.ascontiguousarray(v, np.int32))
        for n in ("u_click", "u_conv", "cost", "drift", "comp_f64"):
            v = getattr(self, n)
            if v is not None:
                setattr(self, n, np.ascontiguousarray(v, np.float64))
        return self

    def c_struct(self) -> _Tape:
        self.normalise()
        t = _Tape()
        for n, _ in _Tape._fields_:
            setattr(t, n, _p(getattr(self, n)))
        return t

    @staticmethod
    def from_lists(volume, comp, u_click, u_conv, rev, impr=None, cost=None, drift=None) -> "Tape":
        """Build from per-keyword python lists/arrays."""
        def csr(lists, dtype):
            off = np.zeros(len(lists) + 1, np.int64)
            off[1:] = np.cumsum([len(x) for x in lists])
            flat = (np.concatenate([np.asarray(x, dtype=dtype).ravel() for x in lists])
                    if len(lists) and off[-1] > 0 else np.zeros(0, dtype))
            return off, np.ascontiguousarray(flat, dtype)
        co, cc = csr(comp, np.int32)
        ko, ku = csr(u_click, np.float64)
        vo, vu = csr(u_conv, np.float64)
        ro, rc = csr(rev, np.int32)
        t = Tape(np.asarray(volume, np.int32), co, cc, ko, ku, vo, vu, ro, rc, drift=drift)
        if impr is not None:
            t.impr = np.asarray(impr, np.int32)
            t.cost_off, t.cost = csr(cost, np.float64)
        return t.normalise()


def _alloc_result(K: int, lanes: bool):
    arrs = dict(
        impressions=np.zeros(K, np.int32), clicks=np.zeros(K, np.int32),
        conversions=np.zeros(K, np.int32), cost=np.zeros(K), revenue=np.zeros(K),
        profit=np.zeros(K), cost_cents=np.zeros(K, np.int64), revenue_cents=np.zeros(K, np.int64))
    if lanes:
        for n in ("lane_I", "lane_B", "lane_S"):
            arrs[n] = np.zeros((SUBSTEPS, K), np.int32)
    r = _Result()
    for n, _ in _Result._fields_[:11]:
        setattr(r, n, _p(arrs.get(n)))
    return r, arrs


def _finish(r: _Result, arrs: Dict[str, np.ndarray]) -> Dict[str, object]:
    out = dict(arrs)
    out["reward"] = float(r.reward)
    out["remaining_budget"] = float(r.remaining_budget)
    out["lanes_run"] = int(r.lanes_run)
    return out


def step_replay(kw: KeywordSet, bid_cents, budget: float, tape: Tape, lanes: bool = True,
                budget_alias: bool = False):
    bc = np.ascontiguousarray(bid_cents, np.int32)
    ks, ts = kw.c_struct(), tape.c_struct()
    r, arrs = _alloc_result(kw.K, lanes)
    rc = lib().orc_step_replay(C.byref(ks), C.c_void_p(bc.ctypes.data), C.c_double(budget),
                               C.c_int(int(budget_alias)), C.byref(ts), C.byref(r))
    if rc:
        raise RuntimeError(f"orc_step_replay failed rc={rc}")
    return _finish(r, arrs)


def step_philox(kw: KeywordSet, bid_cents, budget: float, seed: int, env_id: int, step: int,
                agent: int = 0, lanes: bool = True, record_cap: int = 0,
                budget_alias: bool = False):
    """Free-running step; with record_cap>0 also returns the consumed tape."""
    bc = np.ascontiguousarray(bid_cents, np.int32)
    ks = kw.c_struct()
    r, arrs = _alloc_result(kw.K, lanes)
    rec_ptr = None
    bufs = None
    if record_cap > 0:
        K, cap = kw.K, int(record_cap)
        bufs = dict(
            volume=np.zeros(K, np.int32), comp_cents=np.zeros((K, cap), np.int32),
            u_click=np.zeros((K, cap)), u_conv=np.zeros((K, cap)),
            rev_cents=np.zeros((K, cap), np.int32), impr=np.zeros((K, SUBSTEPS), np.int32),
            cost=np.zeros((K, cap)), comp_f64=np.zeros((K, cap)),
            n_comp=np.zeros(K, np.int32), n_click=np.zeros(K, np.int32),
            n_conv=np.zeros(K, np.int32), n_rev=np.zeros(K, np.int32), n_cost=np.zeros(K, np.int32))
        rec = _Record()
        rec.cap_per_kw = cap
        for n in bufs:
            setattr(rec, n, bufs[n].ctypes.data)
        rec_ptr = C.byref(rec)
    rc = lib().orc_step_philox(C.byref(ks), C.c_void_p(bc.ctypes.data), C.c_double(budget),
                               C.c_int(int(budget_alias)), C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(step),
                               C.c_uint32(agent), C.byref(r), rec_ptr)
    if rc:
        raise RuntimeError(f"orc_step_philox failed rc={rc}")
    out = _finish(r, arrs)
    if bufs is not None:
        K = kw.K
        for n in ("n_comp", "n_click", "n_conv", "n_rev", "n_cost"):
            if bufs[n].max(initial=0) >= record_cap:
                raise RuntimeError("record_cap too small")
        tape = Tape.from_lists(
            bufs["volume"],
            [bufs["comp_cents"][k, :bufs["n_comp"][k]] for k in range(K)],
            [bufs["u_click"][k, :bufs["n_click"][k]] for k in range(K)],
            [bufs["u_conv"][k, :bufs["n_conv"][k]] for k in range(K)],
            [bufs["rev_cents"][k, :bufs["n_rev"][k]] for k in range(K)],
            impr=bufs["impr"] if kw.kind == EXPLICIT else None,
            cost=[bufs["cost"][k, :bufs["n_cost"][k]] for k in range(K)] if kw.kind == EXPLICIT else None)
        if kw.kind == IMPLICIT_MULTI:  # bidders per lane ride in `impr`, the auctions' highest bids in comp_f64
            tape.impr = bufs["impr"].copy()
            tape.comp_f64 = np.concatenate([bufs["comp_f64"][k, :bufs["n_comp"][k]] for k in range(K)] + [np.zeros(0)])
        out["tape"] = tape
    return out


def step_philox_shared(kw: KeywordSet, bid_cents, floor_cents, budget: float, seed: int, world_id: int,
                       step: int, budget_alias: bool = False):
    """One bidder of a shared-auction world: clearing price = max(competitor, highest rival bid)."""
    bc = np.ascontiguousarray(bid_cents, np.int32)
    fc = np.ascontiguousarray(floor_cents, np.int32)
    ks = kw.c_struct()
    r, arrs = _alloc_result(kw.K, True)
    rc = lib().orc_step_philox_shared(C.byref(ks), C.c_void_p(bc.ctypes.data), C.c_void_p(fc.ctypes.data),
                                      C.c_double(budget), C.c_int(int(budget_alias)), C.c_uint64(seed),
                                      C.c_uint32(world_id), C.c_uint32(step), C.byref(r))
    if rc:
        raise RuntimeError(f"orc_step_philox_shared failed rc={rc}")
    return _finish(r, arrs)


def drift_philox(K: int, seed: int, env_id: int, step: int, mag=(0.03, 0.03, 0.03)) -> np.ndarray:
    coeff = np.zeros((3, K))
    m = (C.c_double * 3)(*mag)
    lib().orc_drift_philox(C.c_int32(K), C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(step),
                           m, C.c_void_p(coeff.ctypes.data))
    return coeff


def drift_apply(kw: KeywordSet, mask: np.ndarray, coeff: np.ndarray, init_std: np.ndarray) -> None:
    mask = np.ascontiguousarray(mask, np.uint8)
    coeff = np.ascontiguousarray(coeff, np.float64)
    init_std = np.ascontiguousarray(init_std, np.float64)
    lib().orc_drift_apply(C.c_int32(kw.K), C.c_void_p(mask.ctypes.data), C.c_int32(int(mask.sum())),
                          C.c_void_p(coeff.ctypes.data), C.c_void_p(init_std.ctypes.data),
                          C.c_void_p(kw.vol_mean.ctypes.data), C.c_void_p(kw.ctr.ctypes.data),
                          C.c_void_p(kw.cvr.ctypes.data))


def philox(ctr, key) -> np.ndarray:
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return np.array(list(o), dtype=np.uint32)


def draw4(seed: int, env: int, step: int, agent: int, kw: int, stream: int, idx: int) -> np.ndarray:
    return philox([idx, step, (stream << 28) | (agent << 20) | kw, env],
                  [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])


class BatchOracle:
    """E independent envs advanced by the C oracle in free-running (Philox) mode.

    Mirrors the batched GPU env (same seeds => same integers); also the CPU baseline.
    params: dict name -> [K] (shared keyword set) or [E,K] (per-env sets, needed for drift).
    """

    def __init__(self, kind: int, E: int, K: int, params: Dict[str, np.ndarray], *, seed: int,
                 budget: float = 1000.0, max_days: int = 60, loss_threshold: float = 10000.0,
                 drift_mask: Optional[np.ndarray] = None, drift_mag=(0.03, 0.03, 0.03),
                 env_base: int = 0, step0: int = 0, impression_thresh: float = 0.05,
                 budget_alias: bool = False):
        self.E, self.K, self.kind = E, K, kind
        per_env = drift_mask is not None or np.asarray(params["vol_mean"]).ndim == 2
        self.p = {}
        for n in PARAM_NAMES:
            a = np.asarray(params[n], np.float64)
            if per_env and a.ndim == 1:
                a = np.broadcast_to(a, (E, K))
            self.p[n] = np.ascontiguousarray(a).copy()
        self.mask = None if drift_mask is None else np.ascontiguousarray(drift_mask, np.uint8)
        self.budget = np.full(E, float(budget))
        self.cum_profit = np.zeros(E)
        self.day = np.zeros(E, np.int32)
        b = _Batch()
        b.kind, b.E, b.K = kind, E, K
        b.param_env_stride = K if per_env else 0
        for n in PARAM_NAMES:
            setattr(b, n, self.p[n].ctypes.data)
        for n in ("max_bidders", "participation"):  # IMPLICIT_MULTI keywords
            if params.get(n) is not None:
                a = np.asarray(params[n], np.float64)
                if per_env and a.ndim == 1:
                    a = np.broadcast_to(a, (E, K))
                self.p[n] = np.ascontiguousarray(a).copy()
                setattr(b, n, self.p[n].ctypes.data)
        b.impression_thresh = impression_thresh
        b.drift_mask = _p(self.mask)
        b.drift_mag = (C.c_double * 3)(*drift_mag)
        b.budget, b.cum_profit, b.day = self.budget.ctypes.data, self.cum_profit.ctypes.data, self.day.ctypes.data
        b.max_days, b.loss_threshold = max_days, loss_threshold
        b.seed, b.env_base, b.step = seed, env_base, step0
        b.budget_alias = int(budget_alias)
        self._b = b
        self.out = dict(
            impressions=np.zeros((E, K), np.int32), clicks=np.zeros((E, K), np.int32),
            conversions=np.zeros((E, K), np.int32), cost=np.zeros((E, K)), revenue=np.zeros((E, K)),
            reward=np.zeros(E), terminated=np.zeros(E, np.uint8), truncated=np.zeros(E, np.uint8))

    @property
    def step_count(self) -> int:
        return int(self._b.step)

    def step(self, bids: np.ndarray, n_threads: int = 1) -> Dict[str, np.ndarray]:
        bids = np.ascontiguousarray(np.broadcast_to(bids, (self.E, self.K)), np.float64)
        o = self.out
        rc = lib().orc_batch_step(
            C.byref(self._b), C.c_void_p(bids.ctypes.data), *[C.c_void_p(o[n].ctypes.data) for n in (
                "impressions", "clicks", "conversions", "cost", "revenue", "reward",
                "terminated", "truncated")], C.c_int(n_threads))
        if rc:
            raise RuntimeError(f"orc_batch_step failed rc={rc}")
        return o
