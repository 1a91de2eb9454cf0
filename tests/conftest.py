import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference")


@pytest.fixture(scope="session", autouse=True)
def _cuda_library_present():
    """The in-tree CUDA library normally travels with the tree (built by __graft_entry__.build());
    if it is missing where nvcc exists, build it once so the product can load.  Never a fallback:
    without the library the GPU tests fail loudly in _capi.load()."""
    from adcraft_b200 import build as b
    if not os.path.exists(b.LIB_PATH):
        try:
            b.build()
        except Exception as exc:  # noqa: BLE001
            print(f"conftest: could not build {b.LIB_PATH}: {exc}", file=sys.stderr)
    yield


@pytest.fixture(scope="session")
def orc():
    """The C oracle (test infrastructure)."""
    from oracle import oracle as o
    o.build()
    o.lib()
    return o


def make_implicit_table(rng, K, vol, E=None):
    """Random implicit keyword parameters in the experiment configs' ranges
    (experiment_quantiles.py:16-25)."""
    from adcraft_b200 import keywords as kwm
    shape = (K,) if E is None else (E, K)
    loc = rng.uniform(0.3, 1.0, shape)
    return kwm.KeywordTable(
        kwm.IMPLICIT, np.full(shape, float(vol)), np.floor(1 + rng.random(shape) * 0.5 * vol), loc,
        np.maximum(0.01, rng.uniform(0.01, 0.3, shape) * loc), rng.uniform(0.1, 0.9, shape),
        rng.uniform(0.1, 0.9, shape), rng.uniform(0.3, 1.5, shape),
        np.maximum(0.01, rng.uniform(0.01, 0.3, shape)))


def make_explicit_table(rng, K, E=None):
    from adcraft_b200 import keywords as kwm
    shape = (K,) if E is None else (E, K)
    vm = np.floor(rng.uniform(14, 30, shape))
    mr = rng.beta(2, 5, shape) * 1.5
    return kwm.KeywordTable(
        kwm.EXPLICIT, vm, rng.random(shape) * 0.5 * (vm + 1), rng.random(shape) * 1.5,
        rng.beta(5, 5, shape) * 25, rng.beta(2, 5, shape), rng.beta(5, 2, shape), mr,
        rng.beta(2, 5, shape) * mr)


def oracle_keywordset(orc, table, e=0):
    t = table.env(e)
    from adcraft_b200 import keywords as kwm
    return orc.KeywordSet(t.kind, *[getattr(t, n).copy() for n in kwm.PARAM_NAMES],
                          impression_thresh=t.impression_thresh)
