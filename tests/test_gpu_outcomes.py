"""info["bidding_outcomes"] of the E = 1 adapter against the reference's own string.

The goldens hold what the unmodified reference printed for every step
(``rust.repr_outcomes_py(bidding_outcomes)``, src/lib.rs:250-275, on the per-keyword
``combine_outcomes`` results, bidding_simulation.py:124-147): every click's cost, the per-click
revenues, the impression share with its zero-impression-lane quirk, and the profit accumulated
lane by lane.  In replay mode the adapter must reproduce that string byte for byte."""
import glob
import os

import numpy as np
import pytest

import golden_io

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [c for c in sorted(glob.glob(os.path.join(HERE, "golden", "[rp][eh][cx]_*.npz")))
         if not c.endswith("notebook_lane.npz")]
IDS = [os.path.basename(c)[:-4] for c in CASES]


@pytest.mark.parametrize("path", CASES, ids=IDS)
def test_bidding_outcomes_string_equals_reference(path):
    from adcraft_b200 import keywords as kwm
    from adcraft_b200.gymnasium_kw_env import BiddingSimulation
    from adcraft_b200.tape import DeviceTape
    case = golden_io.load_case(path)
    s0 = case.steps[0]
    assert s0.info_outcomes is not None, "regenerate the goldens (tests/golden/make_golden.py)"
    table = kwm.KeywordTable(case.kind, *[s0.kw_before[n] for n in golden_io.PARAMS], **s0.kw_extra)
    mask = case.meta.get("mask")
    env = BiddingSimulation(num_keywords=case.K, budget=s0.budget, max_days=case.meta.get("max_days", 60),
                            updater_mask=None if mask is None else [bool(m) for m in mask], keywords=table)
    env.reset()
    for i, s in enumerate(case.steps):
        tape = DeviceTape.from_host([s.tape], "cuda")
        budget = np.array([s.budget]) if s.budget_alias else float(s.budget)
        bids = s.bid_cents / 100.0
        if case.meta.get("f32_bids"):
            bids = (s.bid_cents.astype(np.float32) / np.float32(100)).astype(np.float32)
        obs, reward, term, trunc, info = env.step({"keyword_bids": bids, "budget": budget}, tape=tape)
        assert info["bidding_outcomes"] == s.info_outcomes, (i, _first_diff(info["bidding_outcomes"], s.info_outcomes))
        assert np.array_equal(obs["impressions"], s.impressions)
        # the adapter restates the reference's float sums in its order: bit-identical
        assert reward == s.reward
        np.testing.assert_array_equal(obs["cost"], s.cost)
        np.testing.assert_array_equal(obs["revenue"], s.revenue)
        assert float(obs["cumulative_profit"][0]) == s.cumulative_profit


def _first_diff(a, b):
    n = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
    return a[max(0, n - 60):n + 60], b[max(0, n - 60):n + 60]
