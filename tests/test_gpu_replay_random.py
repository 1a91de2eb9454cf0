"""GPU replay (tape mode) against the C oracle on random batched tapes, incl. ragged and empty
streams, binding budgets and the ndarray-aliasing double charge."""
import numpy as np
import pytest

from conftest import make_explicit_table, make_implicit_table, oracle_keywordset

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _random_tape(orc, rng, kind, K, vols, explicit_p=0.5):
    comp, ucl, ucv, rev, cost = [], [], [], [], []
    impr = np.zeros((K, 24), np.int32)
    for k in range(K):
        V = int(vols[k])
        comp.append(rng.integers(0, 160, V))
        if kind == orc.EXPLICIT:
            q = V // 24
            n = [V - 23 * q] + [q] * 23
            impr[k] = [rng.binomial(nn, explicit_p) for nn in n]
            slots = int(np.maximum(impr[k], 1).sum())
            cost.append(np.clip(rng.normal(2.4, 0.3, int(impr[k].sum())), 0, 4.4))
        else:
            slots = V
        ucl.append(rng.random(slots))
        ucv.append(rng.random(slots))
        rev.append(rng.integers(1, 300, slots))
    return orc.Tape.from_lists(vols, comp, ucl, ucv, rev,
                               impr=impr if kind == orc.EXPLICIT else None,
                               cost=cost if kind == orc.EXPLICIT else None)


@pytest.mark.parametrize("n_lanes", [0, 1, "packed"])
@pytest.mark.parametrize("kind_name", ["implicit", "explicit"])
@pytest.mark.parametrize("alias", [False, True])
def test_replay_matches_oracle(orc, kind_name, alias, n_lanes):
    from adcraft_b200.tape import DeviceTape
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(17 + alias)
    kind = orc.IMPLICIT if kind_name == "implicit" else orc.EXPLICIT
    packed = n_lanes == "packed"
    if packed:
        n_lanes = 0
        if kind == orc.EXPLICIT:
            pytest.skip("packed records exist for implicit keywords only")
    K, E = 11, 37
    table = make_implicit_table(rng, K, 60) if kind == orc.IMPLICIT else make_explicit_table(rng, K)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1000.0, device="cuda",
                                  budget_alias=alias, obs_dtype=torch.float64, autoreset=False, n_lanes=n_lanes)
    env.reset()
    budgets = rng.choice([0.0, 0.4, 2.5, 9.0, 30.0, 1e6], size=E)
    cum = np.zeros(E)
    for step in range(3):
        vols = rng.integers(0, 90, (E, K))
        vols[rng.random((E, K)) < 0.1] += 300  # more than one 128-auction trip
        vols[rng.random((E, K)) < 0.15] = 0  # empty units
        tapes = [_random_tape(orc, rng, kind, K, vols[e]) for e in range(E)]
        bids = np.round(rng.uniform(0.01, 1.6, (E, K)), 2)
        obs, reward, term, trunc, _ = env.step_replay(
            {"keyword_bids": torch.from_numpy(bids).cuda(), "budget": torch.from_numpy(budgets).cuda()},
            DeviceTape.from_host(tapes, "cuda", pack=packed))
        for e in range(E):
            kw = oracle_keywordset(orc, table)
            out = orc.step_replay(kw, np.rint(bids[e] * 100).astype(np.int32), float(budgets[e]), tapes[e],
                                  budget_alias=alias)
            for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                         ("sellside_conversions", "conversions")):
                assert np.array_equal(obs[a][e].cpu().numpy(), out[b]), (a, e, step)
            np.testing.assert_allclose(obs["cost"][e].cpu().numpy(), out["cost"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(obs["revenue"][e].cpu().numpy(), out["revenue"], rtol=1e-9, atol=1e-12)
            assert abs(float(reward[e]) - out["reward"]) < 1e-9 * (1 + out["cost"].sum() + out["revenue"].sum())
            rb = float(env._out["remaining_budget"][e])
            assert abs(rb - out["remaining_budget"]) < 1e-9 * (1 + abs(budgets[e]))
            cum[e] += out["reward"]
        np.testing.assert_allclose(obs["cumulative_profit"][:, 0].cpu().numpy(), cum, rtol=1e-9, atol=1e-9)


def _oracle_check(orc, table, obs, reward, bids, budgets, tapes, alias=False):
    for e in range(len(tapes)):
        kw = oracle_keywordset(orc, table)
        out = orc.step_replay(kw, np.rint(bids[e] * 100).astype(np.int32), float(budgets[e]), tapes[e],
                              budget_alias=alias)
        for a, b in (("impressions", "impressions"), ("buyside_clicks", "clicks"),
                     ("sellside_conversions", "conversions")):
            assert np.array_equal(obs[a][e].cpu().numpy(), out[b]), (a, e)
        np.testing.assert_allclose(obs["cost"][e].cpu().numpy(), out["cost"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(obs["revenue"][e].cpu().numpy(), out["revenue"], rtol=1e-9, atol=1e-12)
        assert abs(float(reward[e]) - out["reward"]) < 1e-9 * (1 + out["cost"].sum() + out["revenue"].sum())


def test_packed_replay_edge_records(orc):
    """Packed kernel routes: records larger than a shared-memory stage (generic walk from global
    memory), bids above the 16-bit fast range, negative competitor bids and revenues >= 65536
    cents (generic re-walk of the staged record), and the unit mix of a batch boundary (E*K not a
    multiple of 32 or 8).  Tapes shorter than the walk (overrun -> serial kernel) are covered by the
    budget-bound goldens in test_gpu_golden.py."""
    from adcraft_b200.tape import DeviceTape
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(5)
    K, E = 13, 21
    table = make_implicit_table(rng, K, 60)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda",
                                  obs_dtype=torch.float64, autoreset=False)
    env.reset()
    budgets = np.full(E, 1e9)
    for step in range(2):
        vols = rng.integers(0, 130, (E, K))
        vols[rng.random((E, K)) < 0.1] = 0
        vols[rng.random((E, K)) < 0.1] += 2000          # far beyond one stage
        tapes = []
        for e in range(E):
            comp, ucl, ucv, rev = [], [], [], []
            for k in range(K):
                V = int(vols[e, k])
                c = rng.integers(0, 160, V)
                r = rng.integers(1, 300, V)
                if (e + k) % 5 == 0 and V:
                    c[rng.integers(0, V)] = -rng.integers(1, 100000)   # negative competitor bid
                if (e + k) % 7 == 0 and V:
                    r[: max(1, V // 3)] = rng.integers(65536, 2_000_000_000, max(1, V // 3))
                comp.append(c); ucl.append(rng.random(V)); ucv.append(rng.random(V)); rev.append(r)
            tapes.append(orc.Tape.from_lists(vols[e], comp, ucl, ucv, rev))
        bids = np.round(rng.uniform(0.01, 1.6, (E, K)), 2)
        bids[rng.random((E, K)) < 0.1] = 700.0           # above kMaxFlatBidCents
        tape = DeviceTape.from_host(tapes, "cuda", pack=True)
        obs, reward, *_ = env.step_replay(
            {"keyword_bids": torch.from_numpy(bids).cuda(), "budget": torch.from_numpy(budgets).cuda()}, tape)
        _oracle_check(orc, table, obs, reward, bids, budgets, tapes)


def test_packed_replay_rejects_corrupt_headers(orc):
    """A record whose header claims more entries than the record holds must neither fault nor be
    trusted: the unit flags an overrun and its env is re-walked by the serial kernel from the CSR
    streams, so the step still equals the oracle."""
    from adcraft_b200.tape import DeviceTape
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(6)
    K, E = 5, 9
    table = make_implicit_table(rng, K, 60)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda",
                                  obs_dtype=torch.float64, autoreset=False)
    env.reset()
    vols = rng.integers(1, 100, (E, K))
    tapes = [_random_tape(orc, rng, orc.IMPLICIT, K, vols[e]) for e in range(E)]
    tape = DeviceTape.from_host(tapes, "cuda", pack=True)
    b32 = tape.packed.view(torch.int32)
    off = tape.packed_off.cpu().numpy()
    for u, (field, val) in {3: (2, 1 << 28), 11: (1, 1 << 20), 17: (4, -7), 23: (3, 100000)}.items():
        b32[off[u] // 4 + field] = val
    bids = np.round(rng.uniform(0.01, 1.6, (E, K)), 2)
    budgets = np.full(E, 1e9)
    obs, reward, *_ = env.step_replay(
        {"keyword_bids": torch.from_numpy(bids).cuda(), "budget": torch.from_numpy(budgets).cuda()}, tape)
    _oracle_check(orc, table, obs, reward, bids, budgets, tapes)


@pytest.mark.parametrize("vol", [128, 16])
def test_packed_equals_csr_at_scale(vol):
    """25 600 units with bench-like volumes: the packed kernel (bulk copies through the per-warp
    shared-memory ring, many batches per warp, ring wrap-arounds) must equal the CSR kernel bit
    for bit -- the CSR kernel itself is pinned to the oracle above."""
    from adcraft_b200.tape import DeviceTape
    from adcraft_b200.vector_env import VectorBiddingSimulation
    rng = np.random.default_rng(11)
    K, E = 100, 256
    table = make_implicit_table(rng, K, vol)
    env = VectorBiddingSimulation(E, num_keywords=K, keywords=table, budget=1e9, device="cuda",
                                  obs_dtype=torch.float64, autoreset=False)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    sd = torch.tensor(rng.uniform(1, vol / 2, K), device="cuda")
    V = torch.clamp(torch.round(vol + sd * torch.randn(E, K, device="cuda", dtype=torch.float64, generator=g)), min=0)
    V = V.to(torch.int32)
    V[::7, ::5] = 0
    V[3::50, 1::9] += 700                                   # records larger than the ring
    off = torch.zeros(E * K + 1, dtype=torch.int64, device="cuda")
    off[1:] = torch.cumsum(V.reshape(-1).to(torch.int64), 0)
    n = int(off[-1])
    comp = torch.randint(0, 160, (n,), device="cuda", generator=g, dtype=torch.int32)
    rev = torch.randint(1, 400, (n,), device="cuda", generator=g, dtype=torch.int32)
    loose = DeviceTape(V, off, comp, off, torch.rand(n, device="cuda", dtype=torch.float64, generator=g), off,
                       torch.rand(n, device="cuda", dtype=torch.float64, generator=g), off, rev)
    bids = torch.from_numpy(np.round(rng.uniform(0.2, 1.5, (E, K)), 2)).cuda()
    action = {"keyword_bids": bids}
    keys = ("impressions", "buyside_clicks", "sellside_conversions", "cost", "revenue")
    ref = {k: v.clone() for k, v in env.step_replay(action, loose)[0].items() if k in keys}
    assert int(ref["impressions"].sum()) > 0
    exact = loose.trimmed(ref["impressions"], ref["buyside_clicks"], ref["sellside_conversions"]).pack()
    for tape in (exact, loose.pack()):
        for _ in range(2):                                   # twice: ring / barrier state is per launch
            obs = env.step_replay(action, tape)[0]
            for k in keys:
                assert torch.equal(obs[k], ref[k]), k
