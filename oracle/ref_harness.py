"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference Python.

Nothing under ``adcraft_b200/`` may import this file.  It exists so that
``tests/`` and ``tests/golden/make_golden.py`` can execute the reference's own
``adcraft/bidding_simulation.py``, ``synthetic_kw_classes.py``,
``synthetic_kw_helpers.py``, ``gymnasium_kw_env.py`` and ``gymnasium_kw_utils.py``
straight from ``/root/reference`` (read-only, never copied) in the build
container, where the reference's compiled/3rd-party dependencies are missing:

* ``adcraft.rust`` (PyO3 module built from ``src/lib.rs``; no cargo/rustc here)
  is replaced by :class:`RustShim`, a restatement of the 9 hot-path helpers
  (``src/lib.rs:53-76,92-140,245-275,314-325``).  Its three *random* helpers
  are unseedable in the reference (``thread_rng``), so the shim draws them from
  a pluggable source: a numpy Generator (record mode) or a tape (replay mode).
* ``gymnasium`` is replaced by a ~60 line stub (``Env.reset`` seeding is
  ``np.random.Generator(PCG64(SeedSequence(seed)))`` exactly as gymnasium's
  ``seeding.np_random``; ``spaces.Box/Dict`` only need ``sample``/``contains``).
* ``matplotlib.pyplot`` is an empty module (plots are out of scope).

Two RNG front-ends are provided for the reference's ``np.random.Generator``:

* :class:`RecordingRNG` forwards to a real Generator and logs every call, so a
  reference run can be turned into a replay tape (consumption order).
* :class:`TapeRNG` hands out pre-drawn values from a tape; ``laplace`` returns
  the already abs/rounded competitor bids and ``normal`` the already
  clipped/rounded revenues -- both transforms in the reference
  (``synthetic_kw_helpers.py:66-70,104-113``) are idempotent on such values.

``/root/reference`` does not exist on the GPU box: callers must check
:func:`reference_available` and skip.
"""
from __future__ import annotations

import importlib
import math
import os
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("ADCRAFT_REFERENCE_ROOT", "/root/reference")
# git-ignored copy of the reference's hot-path .py files made by __graft_entry__.build() in the build
# container: it travels to the GPU box with the tree, /root/reference does not
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
HOT_PATH_FILES = (
    "adcraft/__init__.py", "adcraft/bidding_simulation.py", "adcraft/gymnasium_kw_env.py",
    "adcraft/gymnasium_kw_utils.py", "adcraft/synthetic_kw_classes.py", "adcraft/synthetic_kw_helpers.py",
    "adcraft/experiment_utils/__init__.py", "adcraft/experiment_utils/experiment_quantiles.py",
    "adcraft/experiment_utils/experiment_metrics.py", "adcraft/experiment_utils/experiment_configs.py",
    "adcraft/pull_quantiles_data/__init__.py", "adcraft/pull_quantiles_data/quantiles_to_keywords.py",
)


def reference_available(root: Optional[str] = None) -> bool:
    return os.path.isfile(os.path.join(root or REFERENCE_ROOT, "adcraft", "bidding_simulation.py"))


def stage_reference(dst: str = STAGED_ROOT, src: Optional[str] = None) -> Optional[str]:
    """Copy the reference's unmodified hot-path Python files to the git-ignored ``baseline/_ref`` so
    that bench.py's reference arm can time them on the GPU box's host cores.  Returns the staged
    root, or None when the reference tree is not present (on the GPU box: use what travelled)."""
    import shutil
    src = src or REFERENCE_ROOT
    if not reference_available(src):
        return dst if reference_available(dst) else None
    for rel in HOT_PATH_FILES:
        a, b = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
            shutil.copyfile(a, b)
    return dst


# --------------------------------------------------------------------------- #
# RNG front-ends
# --------------------------------------------------------------------------- #
class RecordingRNG(np.random.Generator):
    """``np.random.Generator`` subclass that logs the hot-path draws.

    Subclassing keeps ``isinstance(rng, np.random.Generator)`` true, which the
    reference's ``Keyword._validate_rng`` (synthetic_kw_classes.py:268-277) checks.
    """

    def __init__(self, bit_generator):
        super().__init__(bit_generator)
        self.log: Optional[List[tuple]] = None  # set to a list to start recording

    def _rec(self, name, args, out):
        if self.log is not None:
            self.log.append((name, args, np.array(out, copy=True)))
        return out

    def laplace(self, loc=0.0, scale=1.0, size=None):
        return self._rec("laplace", (loc, scale, size), super().laplace(loc, scale, size))

    def random(self, size=None, *a, **k):
        return self._rec("random", (size,), super().random(size, *a, **k))

    def normal(self, loc=0.0, scale=1.0, size=None):
        return self._rec("normal", (loc, scale, size), super().normal(loc, scale, size))

    def uniform(self, low=0.0, high=1.0, size=None):
        return self._rec("uniform", (low, high, size), super().uniform(low, high, size))

    def binomial(self, n, p, size=None):
        return self._rec("binomial", (n, p, size), super().binomial(n, p, size))


class TapeRNG(np.random.Generator):
    """Replay front-end for one keyword (or for the env-level drift draws).

    The reference consumes, per executed lane of an ImplicitKeyword
    (``bidding_simulation.py:86-111``): ``laplace(size=(1,n))`` ->
    ``random((slots,))`` -> ``random((clicks,))`` -> ``normal(size=convs)``.
    An ExplicitKeyword lane skips ``laplace`` (its auction comes from the rust
    shim).  The tape is day-level per keyword: each stream is consumed
    sequentially across the lanes of one env step.
    """

    def __init__(self):
        super().__init__(np.random.PCG64(0))
        self.comp = np.zeros(0)
        self.u_click = np.zeros(0)
        self.u_conv = np.zeros(0)
        self.rev = np.zeros(0)
        self.drift = []  # list of arrays for env-level uniform() calls
        self.armed = False  # keyword constructors probe their samplers (classes:337-339)
        self.bidders = None  # multi-bidder keywords: bidder count per lane, in lane order
        self.reset_cursors()

    def load(self, comp, u_click, u_conv, rev):
        self.armed = True
        self.comp = np.asarray(comp, dtype=np.float64)
        self.u_click = np.asarray(u_click, dtype=np.float64)
        self.u_conv = np.asarray(u_conv, dtype=np.float64)
        self.rev = np.asarray(rev, dtype=np.float64)
        self.reset_cursors()

    def reset_cursors(self):
        self.i_comp = self.i_click = self.i_conv = self.i_rev = self.i_lane = 0
        self._phase = 0  # 0: next random() is click, 1: next random() is conv
        self.i_drift = 0

    def _take(self, arr, pos, n, what):
        if pos + n > len(arr):
            raise AssertionError(f"tape exhausted: {what} needs {pos + n} > {len(arr)}")
        return arr[pos:pos + n].copy()

    def binomial(self, n, p, size=None):
        """Bidders of the next lane (default ImplicitKeyword, classes:664-665)."""
        m = int(self.bidders[self.i_lane])
        self.i_lane += 1
        return m

    def laplace(self, loc=0.0, scale=1.0, size=None):
        s, n = size
        out = self._take(self.comp, self.i_comp, n, "comp")
        self.i_comp += n
        self._phase = 0
        if self.bidders is None:
            assert s == 1, "single-competitor keywords draw one bid per auction"
            return out.reshape(1, n)
        # multi-bidder keyword: the tape holds each auction's HIGHEST bid; any [s, n] matrix with that
        # column maximum gives the same auction (helpers:116-180 with n=2, num_winners=1)
        return np.vstack([out.reshape(1, n)] + [out.reshape(1, n) - 1.0] * (s - 1)) if s > 0 else np.zeros((0, n))

    def random(self, size=None, *a, **k):
        n = int(size[0]) if isinstance(size, tuple) else int(size)
        if self._phase == 0:
            out = self._take(self.u_click, self.i_click, n, "u_click")
            self.i_click += n
            self._phase = 1
        else:
            out = self._take(self.u_conv, self.i_conv, n, "u_conv")
            self.i_conv += n
            self._phase = 0
        return out

    def normal(self, loc=0.0, scale=1.0, size=None):
        n = int(size)
        if not self.armed:
            return np.full(n, float(loc))
        out = self._take(self.rev, self.i_rev, n, "rev")
        self.i_rev += n
        self._phase = 0
        return out

    def uniform(self, low=0.0, high=1.0, size=None):
        out = np.asarray(self.drift[self.i_drift], dtype=np.float64)
        self.i_drift += 1
        n = int(size[0]) if isinstance(size, tuple) else int(size)
        assert len(out) == n
        return out.copy()


# --------------------------------------------------------------------------- #
# adcraft.rust shim
# --------------------------------------------------------------------------- #
def _round_half_away(x: float) -> int:
    # Rust f64::round (src/lib.rs:324): half away from zero.
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


class RustShim(types.ModuleType):
    """Restatement of the hot-path helpers of ``src/lib.rs`` (see module doc).

    ``source`` decides where the three random helpers get their draws:
    ``("numpy", gen)`` draws fresh values (and logs them into ``self.log`` when
    it is a list); ``("tape", obj)`` asks ``obj.volume(mean, std)``,
    ``obj.binomial(n, p)`` and ``obj.costs(x, n)`` (called in the reference's order).
    """

    def __init__(self):
        super().__init__("adcraft.rust")
        self.source = ("numpy", np.random.default_rng(0xC0FFEE))
        self.log: Optional[List[tuple]] = None

    # -- deterministic helpers ------------------------------------------------
    @staticmethod
    def sigmoid(x, s, t):  # src/lib.rs:290-294
        return 1.0 / (1.0 + math.exp(-s * (x - t)))

    @staticmethod
    def probify_float(x, y, z):  # src/lib.rs:296-300
        return min(max(x, y), z)

    def threshold_sigmoid(self, p, params):  # src/lib.rs:92-105
        thresh_in = float(params["impression_thresh"])
        intercept = float(params["impression_bid_intercept"])
        slope = float(params["impression_slope"])
        halver = 2.0 + 1e-10
        thresh = self.probify_float(halver * thresh_in, 0.0, 1.0) / halver
        r = self.sigmoid(float(p), slope, intercept)
        return self.probify_float((1.0 + 2.0 * thresh) * r - thresh, 0.0, 1.0)

    @staticmethod
    def sum_list(xs):  # src/lib.rs:113-116: sequential left-to-right f64 sum
        acc = 0.0
        for v in xs:
            acc += float(v)
        return acc

    @staticmethod
    def sum_array(arr):  # src/lib.rs:107-111: ndarray::sum (8-way unrolled)
        a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        if a.dtype != np.float64:
            raise TypeError("sum_array expects float64")
        p = [0.0] * 8
        n8 = len(a) // 8 * 8
        for i in range(0, n8, 8):
            for j in range(8):
                p[j] += float(a[i + j])
        acc = 0.0
        acc += p[0] + p[4]
        acc += p[1] + p[5]
        acc += p[2] + p[6]
        acc += p[3] + p[7]
        for v in a[n8:]:
            acc += float(v)
        return acc

    @staticmethod
    def sum_array_bool(arr):  # src/lib.rs:123-127
        return int(np.count_nonzero(np.asarray(arr, dtype=bool)))

    @staticmethod
    def sum_list_bool(xs):  # src/lib.rs:118-121
        return int(sum(1 for b in xs if b))

    @staticmethod
    def list_to_zeros(xs):  # src/lib.rs:136-140
        return np.zeros(len(xs), dtype=np.float64)

    @staticmethod
    def array_to_zeros(arr):  # src/lib.rs:129-134
        return np.zeros(np.asarray(arr).size, dtype=np.float64)

    @staticmethod
    def repr_outcomes_py(outcomes):  # src/lib.rs:250-275 (info string only)
        def f(v):  # Rust `{}` on f64: shortest round-trip digits, always positional, 1.0 -> "1"
            from decimal import Decimal
            v = float(v)
            if v != v or v in (float("inf"), float("-inf")):
                return "NaN" if v != v else ("inf" if v > 0 else "-inf")
            s = format(Decimal(repr(v)), "f")
            return s.rstrip("0").rstrip(".") if "." in s else s

        def g(v):  # Rust `{:?}` on f64: like Python's repr, exponent written 1.5e-7 / 1e16
            s = repr(float(v))
            if "e" in s:
                mant, exp = s.split("e")
                return f"{mant[:-2] if mant.endswith('.0') else mant}e{int(exp)}"
            return s

        def fl(vs):  # Rust `{:?}` on Vec<f64>
            return "[" + ", ".join(g(v) for v in vs) + "]"

        parts = []
        for o in outcomes:
            parts.append(
                "{" + f"'bid': {f(o['bid'])}, 'impressions': {int(o['impressions'])}, "
                f"'impression_share': {f(o['impression_share'])}, "
                f"'buyside_clicks': {int(o['buyside_clicks'])}, 'costs': {fl(o['costs'])}, "
                f"'sellside_conversions': {int(o['sellside_conversions'])}, "
                f"'revenues': {fl(o['revenues'])}, "
                f"'revenues_per_cost': {fl(o['revenues_per_cost'])}, 'profit': {f(o['profit'])}" + "}"
            )
        return "[" + ", ".join(parts) + "]"

    # -- random helpers ---------------------------------------------------------
    def _log(self, name, args, out):
        if self.log is not None:
            self.log.append((name, args, np.array(out, copy=True)))
        return out

    def nonneg_int_normal_sampler(self, the_mean, std):  # src/lib.rs:314-325
        kind, src = self.source
        if kind == "tape":
            out = int(src.volume(float(the_mean), float(std)))
        else:
            raw = float(src.normal(float(the_mean), float(std)))
            out = _round_half_away(max(raw, 0.0))
        return self._log("rust.volume", (float(the_mean), float(std)), out)

    def binomial_impressions(self, n, p):  # src/lib.rs:69-76
        kind, src = self.source
        if not (0.0 <= p <= 1.0):
            raise RuntimeError("Binomial::new(n,p).unwrap() panics for p outside [0,1]")
        if kind == "tape":
            out = int(src.binomial(int(n), float(p)))
        else:
            out = int(src.binomial(int(n), float(p)))
        return self._log("rust.binomial", (int(n), float(p)), out)

    def cost_create(self, x, n):  # src/lib.rs:53-67 (constant 4.4)
        kind, src = self.source
        n = int(n)
        if kind == "tape":
            out = np.array(src.costs(float(x), n), dtype=np.float64)
            assert len(out) == n, "tape exhausted: explicit costs"
        else:
            xs = math.sqrt(float(x))
            noise = src.normal(0.0, 1e-10 + xs / 6.0, size=n)
            out = np.clip((xs / 4.0 + 4.4 / 2.0) + noise, 0.0, 4.4)
        return self._log("rust.cost_create", (float(x), n), out)

    # unused-on-hot-path cost helpers (tests/rust/test_helpers.py)
    def cost_trans(self, arr):  # src/lib.rs:32-51
        _, src = self.source
        p = np.asarray(arr, dtype=np.float64)
        s = np.sqrt(p)
        return np.clip(s / 4.0 + p / 2.0 + src.normal(size=p.shape) * (1e-10 + s / 6.0), 0.0, p)

    def cost_mut(self, arr):  # src/lib.rs:16-30
        arr[...] = self.cost_trans(arr)


# --------------------------------------------------------------------------- #
# gymnasium / matplotlib stubs
# --------------------------------------------------------------------------- #
def _make_gymnasium_stub() -> types.ModuleType:
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.zeros(self.shape, dtype=self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return (x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                    and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high)))

    class Dict(Space):
        def __init__(self, d):
            self.spaces = dict(d)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return set(x.keys()) == set(self.spaces.keys()) and all(
                self.spaces[k].contains(v) for k, v in x.items())

        def __getitem__(self, k):
            return self.spaces[k]

    class Env:
        _np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None or self._np_random is None:
                ss = np.random.SeedSequence(seed)
                self._np_random = RecordingRNG(np.random.PCG64(ss))

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = RecordingRNG(np.random.PCG64(np.random.SeedSequence()))
            return self._np_random

        @np_random.setter
        def np_random(self, v):
            self._np_random = v

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

    spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict
    gym.Env, gym.Wrapper, gym.spaces, gym.Space = Env, Wrapper, spaces, Space
    return gym, spaces


_LOADED: Dict[str, Any] = {}


def load_reference(root: Optional[str] = None) -> Dict[str, Any]:
    """Import the reference modules (once) and return them with the shim.  ``root``: the tree to
    import from (default /root/reference; bench.py passes the staged copy)."""
    global REFERENCE_ROOT
    if _LOADED:
        return _LOADED
    if root is not None:
        REFERENCE_ROOT = root
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "gymnasium" not in sys.modules:
        gym, spaces = _make_gymnasium_stub()
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    shim = RustShim()
    # `from adcraft import rust` resolves the attribute on the package; make the
    # package importable from the read-only tree, then plant the shim.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    pkg = importlib.import_module("adcraft")
    if not getattr(pkg, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("a different `adcraft` package is already imported")
    sys.modules["adcraft.rust"] = shim
    pkg.rust = shim
    _LOADED.update(
        shim=shim,
        helpers=importlib.import_module("adcraft.synthetic_kw_helpers"),
        classes=importlib.import_module("adcraft.synthetic_kw_classes"),
        bsim=importlib.import_module("adcraft.bidding_simulation"),
        utils=importlib.import_module("adcraft.gymnasium_kw_utils"),
        env=importlib.import_module("adcraft.gymnasium_kw_env"),
        quantiles=importlib.import_module("adcraft.experiment_utils.experiment_quantiles"),
        metrics=importlib.import_module("adcraft.experiment_utils.experiment_metrics"),
        q2k=importlib.import_module("adcraft.pull_quantiles_data.quantiles_to_keywords"),
    )
    return _LOADED


def experiment_keyword_config(mean_volume: int, cvr: float, tmpdir: str) -> dict:
    """keyword_config dict of the reference's experiment configs
    (``experiment_utils/experiment_configs.py:15-27``) writing its CSV to tmpdir."""
    ref = load_reference()
    return {
        "outer_directory": tmpdir,
        "mean_volume": mean_volume,
        "conversion_rate": cvr,
        "make_quant_func": ref["quantiles"].make_experiment_quantiles,
        "load_quant_func": ref["quantiles"].load_experiment_quantiles,
    }
