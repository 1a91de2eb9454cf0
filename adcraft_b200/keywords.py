"""Keyword parameter tables and the reset-time keyword factories.

Host-side (numpy) mirror of the reference's keyword sampling, consuming the env's numpy
Generator in exactly the reference's order so that ``reset(seed=s)`` yields the same keywords:

* :func:`sample_random_keywords`  <- ``adcraft/gymnasium_kw_utils.py:113-156`` (ExplicitKeyword)
* :func:`sample_implicit_keywords_from_quantiles` <- ``gymnasium_kw_utils.py:260-349`` +
  ``pull_quantiles_data/quantiles_to_keywords.py:13-28`` + the in-memory equivalent of
  ``experiment_utils/experiment_quantiles.py:7-84`` (the singleton quantile rows are built in
  memory by default; the reference's CSV wire format -- ``count_/min_/median_/max_<param>``
  columns -- is read and written by ``read_quantile_csv`` / ``write_quantile_csv``, with the
  reference's ``quantiles_folder`` default loader and ``make/load_quant_func`` conventions).

The table is SoA float64: ``((vol_mean, vol_std), loc|intercept, scale|slope, bctr, sctr,
mean_rev, std_rev)`` of ``gymnasium_kw_utils.py:20-28``; for implicit keywords ``p2`` holds the
Laplace *scale* (the reference's params tuple stores ``1/scale``, utils:195).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

IMPLICIT, EXPLICIT, IMPLICIT_MULTI = 0, 1, 2
PARAM_NAMES = ("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std")

# experiment_quantiles.py:16-25
GENERIC_SPARSITY_QUANTILES = {
    "vol": [64, 128, 256],
    "ave_cpc": [0.3, 0.55, 1],
    "std_cpc": [0.01, 0.15, 0.3],
    "bctr": [0.1, 0.5, 0.9],
    "sctr": [0.1, 0.5, 0.9],
    "rpsc": [0.3, 1.0, 1.5],
    "std_rpsc": [0.01, 0.15, 0.3],
}


@dataclass
class KeywordTable:
    kind: int
    vol_mean: np.ndarray
    vol_std: np.ndarray
    p1: np.ndarray
    p2: np.ndarray
    ctr: np.ndarray
    cvr: np.ndarray
    rev_mean: np.ndarray
    rev_std: np.ndarray
    impression_thresh: float = 0.05  # gymnasium_kw_utils.py:81
    # IMPLICIT_MULTI (the class-default ImplicitKeyword, synthetic_kw_classes.py:649-688): bidders per
    # lane ~ Binomial(max_bidders, participation); p1 / p2 are the SIGNED Laplace's loc / scale
    max_bidders: Optional[np.ndarray] = None
    participation: Optional[np.ndarray] = None

    def __post_init__(self):
        for n in PARAM_NAMES:
            setattr(self, n, np.ascontiguousarray(getattr(self, n), dtype=np.float64))
        shapes = {getattr(self, n).shape for n in PARAM_NAMES}
        assert len(shapes) == 1, f"inconsistent keyword parameter shapes: {shapes}"
        if self.kind == IMPLICIT_MULTI:
            shape = self.vol_mean.shape
            mb = 30.0 if self.max_bidders is None else self.max_bidders      # classes:659-662
            pr = 3.0 / 5.0 if self.participation is None else self.participation  # classes:663
            self.max_bidders = np.ascontiguousarray(np.broadcast_to(np.asarray(mb, np.float64), shape)).copy()
            self.participation = np.ascontiguousarray(np.broadcast_to(np.asarray(pr, np.float64), shape)).copy()

    @property
    def K(self) -> int:
        return self.vol_mean.shape[-1]

    @property
    def per_env(self) -> bool:
        return self.vol_mean.ndim == 2

    @property
    def env_stride(self) -> int:
        return self.K if self.per_env else 0

    def env(self, e: int) -> "KeywordTable":
        """The keyword set of env e as a shared ([K]) table."""
        if not self.per_env:
            return self
        return KeywordTable(self.kind, *[getattr(self, n)[e] for n in PARAM_NAMES],
                            impression_thresh=self.impression_thresh,
                            max_bidders=None if self.max_bidders is None else self.max_bidders[e],
                            participation=None if self.participation is None else self.participation[e])

    def describe(self, e: int = 0) -> str:
        """``repr_all_params`` (gymnasium_kw_utils.py:352-380) for env e."""
        t = self.env(e)
        names = ["volume", "imp_intercept", "imp_slope", "bctr", "sctr", "mean revenue", "std revenue"]
        rows = []
        for k in range(t.K):
            p2 = t.p2[k] if t.kind == EXPLICIT else 1.0 / t.p2[k]
            vm, vs = t.vol_mean[k], t.vol_std[k]
            if t.kind == IMPLICIT and vm == int(vm) and vs == int(vs):
                vol = (int(vm), int(vs))
            else:
                vol = (float(vm), float(vs))
            vals = [vol, float(t.p1[k]), float(p2), float(t.ctr[k]), float(t.cvr[k]),
                    float(t.rev_mean[k]), float(t.rev_std[k])]
            rows.append(f"kw{k} params:\n " + ",   ".join(f"{n}: {v}" for n, v in zip(names, vals)))
        return "\n".join(rows)


def _stack(tables) -> KeywordTable:
    t0 = tables[0]
    if len(tables) == 1:
        return t0
    return KeywordTable(t0.kind, *[np.stack([getattr(t, n) for t in tables]) for n in PARAM_NAMES],
                        impression_thresh=t0.impression_thresh)


def _consume_constructor_draws(rng: np.random.Generator, K: int) -> None:
    """Keyword.__init__ probes its reward sampler with rds(2), rds(5), rds(5)
    (synthetic_kw_classes.py:337-339), i.e. 12 normals per keyword from the shared Generator.
    They do not influence the parameters; consumed only to leave `rng` where the reference does."""
    for _ in range(K):
        rng.normal(0.0, 1.0, 2)
        rng.normal(0.0, 1.0, 5)
        rng.normal(0.0, 1.0, 5)


def _sample_random_one(K: int, rng: np.random.Generator) -> KeywordTable:
    # gymnasium_kw_utils.py:129-140, same draw order
    v_mean = (2 ** rng.beta(2, 5, size=K) * 15 - 1).astype(int)
    v_std = rng.random(size=K) * 0.5 * (v_mean + 1)
    sctr = rng.beta(5, 2, size=K)
    intercept = rng.random(size=K) * 1.5
    mean_rev = rng.beta(2, 5, size=K) * 1.5
    std_rev = rng.beta(2, 5, size=K) * mean_rev
    bctr = rng.beta(2, 5, size=K)
    slope = rng.beta(5, 5, size=K) * 25
    _consume_constructor_draws(rng, K)
    # Keyword._buyside_ctr_init / _sellside_paid_ctr_init probify the rates (classes:405-407)
    return KeywordTable(EXPLICIT, v_mean, v_std, intercept, slope, np.clip(bctr, 0.0, 1.0),
                        np.clip(sctr, 0.0, 1.0), mean_rev, std_rev)


def sample_random_keywords(num_keywords: int, rng: np.random.Generator,
                           num_envs: Optional[int] = None) -> KeywordTable:
    """ExplicitKeyword parameters (default env); ``num_envs`` draws one set per env in turn."""
    n = 1 if num_envs is None else int(num_envs)
    out = _stack([_sample_random_one(num_keywords, rng) for _ in range(n)])
    if num_envs is not None and not out.per_env:
        out = KeywordTable(out.kind, *[getattr(out, n_)[None] for n_ in PARAM_NAMES],
                           impression_thresh=out.impression_thresh)
    return out


def sample_from_quantiles(n, num_buckets, mins, meds, maxs, rng) -> np.ndarray:
    """quantiles_to_keywords.py:13-28: uniform bucket, then piecewise-linear interpolation."""
    buckets = rng.integers(low=0, high=num_buckets, size=(n,))
    samples = rng.random(size=(n,))
    mins, meds, maxs = (np.asarray(a, dtype=np.float64) for a in (mins, meds, maxs))
    lo, mid, hi = mins[buckets], meds[buckets], maxs[buckets]
    # np.interp(q, [0, .5, 1], [lo, mid, hi]) evaluated per element
    out = np.where(samples <= 0.5, lo + (mid - lo) * (samples / 0.5),
                   mid + (hi - mid) * ((samples - 0.5) / 0.5))
    exact = np.array([np.interp(q, [0.0, 0.5, 1.0], [a, b, c])
                      for q, a, b, c in zip(samples, lo, mid, hi)]) if n <= 4096 else out
    return exact


QUANTILE_PARAMS = ("vol", "ave_cpc", "std_cpc", "bctr", "sctr", "rpsc", "std_rpsc")
_QUANTILE_STATS = ("count", "min", "median", "max")


def write_quantile_csv(cols: Dict[str, np.ndarray], path: str) -> None:
    """Write a quantile table in the reference's wire format: what ``DataFrame.to_csv(path)`` produces
    for columns ``count_/min_/median_/max_<param>`` (experiment_quantiles.py:28-33,68-73) -- an
    unnamed index column first, one row per quantile bucket, empty cell for NaN."""
    names = [f"{st}_{p}" for p in QUANTILE_PARAMS for st in _QUANTILE_STATS if f"{st}_{p}" in cols]
    names += [c for c in cols if c not in names]
    arrs = [np.atleast_1d(np.asarray(cols[c], dtype=np.float64)) for c in names]
    n = len(arrs[0])
    assert all(len(a) == n for a in arrs), "quantile columns must have one value per bucket"

    def cell(v: float) -> str:
        if v != v:
            return ""
        return str(int(v)) if float(v).is_integer() and abs(v) < 2 ** 53 else repr(float(v))

    with open(path, "w", newline="") as f:
        f.write("," + ",".join(names) + "\n")
        for i in range(n):
            f.write(str(i) + "," + ",".join(cell(a[i]) for a in arrs) + "\n")


def read_quantile_csv(path: str) -> Dict[str, np.ndarray]:
    """Read a quantile table written by pandas (``to_csv``) or by ``write_quantile_csv``: float64
    columns ``count_/min_/median_/max_<param>``; the index column and anything else is dropped.
    Parsed by ``pandas.read_csv`` like the reference does (utils:254) when pandas is importable --
    its default float parser is not correctly rounded, so only the same parser gives the same
    last bits -- else by Python's ``float``."""
    try:
        import pandas as pd
    except ImportError:
        pd = None
    if pd is not None:
        data = pd.read_csv(path)
        return {c: np.asarray(data[c], dtype=np.float64) for c in data.columns
                if "_" in c and c.split("_")[0] in _QUANTILE_STATS}
    import csv
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    header, body = rows[0], [r for r in rows[1:] if r]
    out = {}
    for j, name in enumerate(header):
        if "_" in name and name.split("_")[0] in _QUANTILE_STATS:
            out[name] = np.array([float(r[j]) if j < len(r) and r[j] != "" else np.nan for r in body],
                                 dtype=np.float64)
    return out


def make_experiment_quantiles(keyword_config: Dict) -> None:
    """experiment_quantiles.py:37-47,68-73: the singleton quantile bucket of the experiment configs,
    written to ``<outer_directory>/<mean_volume>_<conversion_rate>.csv`` (usable as ``make_quant_func``)."""
    v, cvr = keyword_config["mean_volume"], keyword_config["conversion_rate"]
    cols = quantile_rows_from_config({"mean_volume": v, "conversion_rate": cvr})
    write_quantile_csv(cols, f"{keyword_config['outer_directory']}/{v}_{cvr}.csv")


def load_experiment_quantiles(keyword_config: Dict) -> Dict[str, np.ndarray]:
    """experiment_quantiles.py:76-84 (usable as ``load_quant_func``)."""
    v, cvr = keyword_config["mean_volume"], keyword_config["conversion_rate"]
    return read_quantile_csv(f"{keyword_config['outer_directory']}/{v}_{cvr}.csv")


def load_quantile_rows_from_csv(keyword_config: Dict) -> Optional[Dict[str, np.ndarray]]:
    """gymnasium_kw_utils.py:238-257, the reference's default ``load_quant_func``:
    ``<outer_directory><quantiles_folder>auction_data.csv`` (plain string concatenation, as there),
    None when the file does not exist."""
    outer = keyword_config.get("outer_directory", os.getcwd().replace(os.sep, "/") + "/quantile_dfs/")
    path = outer + keyword_config.get("quantiles_folder") + "auction_data.csv"
    return read_quantile_csv(path) if os.path.isfile(path) else None


def quantile_rows_from_config(keyword_config: Dict) -> Dict[str, np.ndarray]:
    """Quantile table as a dict of columns ``count_/min_/median_/max_<param>``.

    Accepts (a) a user ``load_quant_func`` (+ optional ``make_quant_func``) exactly like the
    reference (utils:281-289), returning a DataFrame or a dict of columns; (b) an explicit
    ``quantile_table`` mapping; (c) ``quantiles_folder`` alone: the reference's default CSV loader
    (utils:238-257); or (d) the experiment configs' ``mean_volume`` / ``conversion_rate`` pair
    (experiment_quantiles.py:37-47) built in memory."""
    load = keyword_config.get("load_quant_func")
    if load is None and "quantile_table" not in keyword_config and keyword_config.get("quantiles_folder", False):
        load = load_quantile_rows_from_csv
    if load is not None:
        if not keyword_config.get("quantiles_folder", False):
            make = keyword_config.get("make_quant_func")
            if make is not None:
                make(keyword_config)
        data = load(keyword_config)
        assert data is not None, "Invalid quantile parameters specified in keyword_config for data"
        names = data.keys() if isinstance(data, dict) else data.columns
        return {c: np.asarray(data[c], dtype=np.float64) for c in names if c.split("_")[0] in _QUANTILE_STATS}
    if "quantile_table" in keyword_config:
        return {k: np.atleast_1d(np.asarray(v, dtype=np.float64))
                for k, v in keyword_config["quantile_table"].items()}
    d = {k: list(v) for k, v in GENERIC_SPARSITY_QUANTILES.items()}
    if "mean_volume" in keyword_config:
        v = keyword_config["mean_volume"]
        d["vol"] = [v, v, v]
    if "conversion_rate" in keyword_config:
        c = keyword_config["conversion_rate"]
        d["sctr"] = [c, c, c]
    if "clickthrough_rate" in keyword_config:
        c = keyword_config["clickthrough_rate"]
        d["bctr"] = [c, c, c]
    cols = {}
    for name, (lo, mid, hi) in d.items():  # singleton_mmm_dict (experiment_quantiles.py:7-14)
        cols[f"count_{name}"] = np.array([3.0])
        cols[f"min_{name}"] = np.array([float(lo)])
        cols[f"median_{name}"] = np.array([float(mid)])
        cols[f"max_{name}"] = np.array([float(hi)])
    return cols


def _sample_implicit_one(K: int, rng: np.random.Generator, data: Dict[str, np.ndarray],
                         no_volume_prob: float) -> KeywordTable:
    nrows = len(data["min_vol"])
    v = sample_from_quantiles(K, nrows, data["min_vol"], data["median_vol"], data["max_vol"], rng)
    vol_mean = np.zeros(K)
    vol_std = np.zeros(K)
    for i in range(K):  # utils:296-309: the condition's draw comes first, then the std's draw
        has_vol = rng.random() > no_volume_prob and not np.isnan(v[i])
        if has_vol:
            vol_mean[i] = int(v[i])
            vol_std[i] = int(1 + rng.random() * 0.5 * v[i])
        else:
            vol_mean[i] = 0
            vol_std[i] = rng.random() * 0.5
    cols = []
    for param in ["ave_cpc", "std_cpc", "bctr", "sctr", "rpsc", "std_rpsc"]:
        keep = data[f"count_{param}"] > 0
        vals = sample_from_quantiles(K, int(keep.sum()), data[f"min_{param}"][keep],
                                     data[f"median_{param}"][keep], data[f"max_{param}"][keep], rng)
        if param.startswith("std_"):  # un-normalise: multiplier on the preceding average
            vals = np.maximum(0.01, vals * cols[-1])
        cols.append(vals)
    loc, scale, bctr, sctr, rev, rev_std = cols
    _consume_constructor_draws(rng, K)
    return KeywordTable(IMPLICIT, vol_mean, vol_std, loc, scale, np.clip(bctr, 0.0, 1.0),
                        np.clip(sctr, 0.0, 1.0), rev, rev_std)


def sample_implicit_keywords_from_quantiles(num_keywords: int, rng: np.random.Generator,
                                            keyword_config: Dict,
                                            num_envs: Optional[int] = None) -> KeywordTable:
    """ImplicitKeyword parameters with one competitor (utils:260-349, :159-195)."""
    data = quantile_rows_from_config(keyword_config)
    p0 = keyword_config.get("no_vol_prob", 0.0)
    n = 1 if num_envs is None else int(num_envs)
    out = _stack([_sample_implicit_one(num_keywords, rng, data, p0) for _ in range(n)])
    if num_envs is not None and not out.per_env:
        out = KeywordTable(out.kind, *[getattr(out, n_)[None] for n_ in PARAM_NAMES],
                           impression_thresh=out.impression_thresh)
    return out


# ----------------------------------------------------------------------------------------------
# device-side sampling: the same distributions drawn with torch's generator on the
# GPU, for per-env keyword sets at sizes where host sampling + upload would dominate reset time.
# Not stream-compatible with numpy (use the host factories above for the reference's exact draws).
# ----------------------------------------------------------------------------------------------
def sample_implicit_keywords_device(num_envs: int, num_keywords: int, keyword_config: Dict, device,
                                    generator=None) -> Dict[str, "object"]:
    """Per-env implicit keyword parameters as float64 CUDA tensors [E, K] (utils:260-349 semantics:
    uniform bucket, piecewise-linear quantile interpolation, std_* as multipliers floored at 0.01,
    vol_std = int(1 + U * 0.5 * vol), zero-volume keywords with probability ``no_vol_prob``)."""
    import torch
    data = quantile_rows_from_config(keyword_config)
    E, K = int(num_envs), int(num_keywords)
    f64 = torch.float64

    def draw(param):
        keep = data[f"count_{param}"] > 0 if f"count_{param}" in data else np.ones(len(data[f"min_{param}"]), bool)
        lo, mid, hi = (torch.tensor(data[f"{q}_{param}"][keep], dtype=f64, device=device)
                       for q in ("min", "median", "max"))
        b = torch.randint(0, len(lo), (E, K), device=device, generator=generator)
        q = torch.rand(E, K, dtype=f64, device=device, generator=generator)
        lo, mid, hi = lo[b], mid[b], hi[b]
        return torch.where(q <= 0.5, lo + (mid - lo) * (q / 0.5), mid + (hi - mid) * ((q - 0.5) / 0.5))

    v = draw("vol")
    has_vol = torch.rand(E, K, dtype=f64, device=device, generator=generator) > keyword_config.get("no_vol_prob", 0.0)
    u = torch.rand(E, K, dtype=f64, device=device, generator=generator)
    vol_mean = torch.where(has_vol, torch.floor(v), torch.zeros_like(v))
    vol_std = torch.where(has_vol, torch.floor(1 + u * 0.5 * v), u * 0.5)
    loc = draw("ave_cpc")
    scale = torch.clamp(draw("std_cpc") * loc, min=0.01)
    bctr = draw("bctr").clamp(0.0, 1.0)
    sctr = draw("sctr").clamp(0.0, 1.0)
    rev = draw("rpsc")
    rev_std = torch.clamp(draw("std_rpsc") * rev, min=0.01)
    return dict(vol_mean=vol_mean, vol_std=vol_std, p1=loc, p2=scale, ctr=bctr, cvr=sctr,
                rev_mean=rev, rev_std=rev_std)


def sample_random_keywords_device(num_envs: int, num_keywords: int, device, generator=None) -> Dict[str, "object"]:
    """Per-env ExplicitKeyword parameters (the default env's factory, gymnasium_kw_utils.py:113-156) as
    float64 CUDA tensors [E, K], drawn on the device: v_mean = int(2 ** Beta(2,5) * 15 - 1),
    v_std = U * 0.5 * (v_mean + 1), sctr ~ Beta(5,2), intercept ~ 1.5 U, mean_rev ~ 1.5 Beta(2,5),
    std_rev ~ Beta(2,5) * mean_rev, bctr ~ Beta(2,5), slope ~ 25 Beta(5,5).  Beta(a, b) with integer
    shapes is drawn exactly as the a-th smallest of a + b - 1 uniforms (order statistics), so the
    torch Generator is the only source of randomness.  Same distributions as the host factory
    (``sample_random_keywords``, which reproduces the reference's draws bit for bit); KS-tested."""
    import torch
    E, K = int(num_envs), int(num_keywords)
    f64 = torch.float64

    def uniform():
        return torch.rand(E, K, dtype=f64, device=device, generator=generator)

    def beta(a: int, b: int):
        u = torch.rand(E, K, a + b - 1, dtype=f64, device=device, generator=generator)
        return torch.kthvalue(u, a, dim=-1).values

    v_mean = torch.trunc(torch.pow(2.0, beta(2, 5)) * 15.0 - 1.0)
    v_std = uniform() * 0.5 * (v_mean + 1.0)
    sctr = beta(5, 2)
    intercept = uniform() * 1.5
    mean_rev = beta(2, 5) * 1.5
    std_rev = beta(2, 5) * mean_rev
    bctr = beta(2, 5)
    slope = beta(5, 5) * 25.0
    return dict(vol_mean=v_mean, vol_std=v_std, p1=intercept, p2=slope, ctr=bctr.clamp(0.0, 1.0),
                cvr=sctr.clamp(0.0, 1.0), rev_mean=mean_rev, rev_std=std_rev)
