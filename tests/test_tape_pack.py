"""The packed replay-tape format (include/adcraft_b200.h, `adc_tape.packed`) on the CPU: pack() is
plain tensor ops, so the record layout the CUDA kernel parses is pinned here without a GPU."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _random_csr(rng, U, max_v):
    V = rng.integers(0, max_v, U)
    V[rng.random(U) < 0.2] = 0
    n_comp = V + rng.integers(0, 3, U)                    # streams may be longer than what is consumed
    n_click = rng.integers(0, max_v, U)
    n_conv = rng.integers(0, max_v // 2 + 1, U)
    n_rev = rng.integers(0, max_v // 2 + 1, U)

    def csr(counts, gen):
        off = np.zeros(U + 1, np.int64)
        off[1:] = np.cumsum(counts)
        return torch.from_numpy(off), torch.from_numpy(gen(int(off[-1])))

    def comps(n):
        c = rng.integers(0, 200, n).astype(np.int32)
        c[rng.random(n) < 0.01] = -5                        # a negative competitor bid: not a narrow record
        return c

    def revs(n):
        r = rng.integers(1, 500, n).astype(np.int32)
        r[rng.random(n) < 0.01] = 70000                     # beyond 16 bits: not a narrow record
        return r

    comp = csr(n_comp, comps)
    click = csr(n_click, lambda n: rng.random(n))
    conv = csr(n_conv, lambda n: rng.random(n))
    rev = csr(n_rev, revs)
    return V, comp, click, conv, rev


def test_packed_records_hold_the_csr_streams():
    from adcraft_b200.tape import DeviceTape
    rng = np.random.default_rng(3)
    E, K = 7, 9
    V, comp, click, conv, rev = _random_csr(rng, E * K, 40)
    t = DeviceTape(torch.from_numpy(V.astype(np.int32)).view(E, K), comp[0], comp[1], click[0], click[1],
                   conv[0], conv[1], rev[0], rev[1]).pack()
    buf, off = t.packed.numpy(), t.packed_off.numpy()
    assert off[0] == 0 and np.all(off % 16 == 0) and np.all(np.diff(off) >= 0) and off[-1] == buf.size
    seen_flags = set()
    for u in range(E * K):
        rec = buf[off[u]:off[u + 1]]
        if V[u] == 0:
            assert rec.size == 0                           # an empty record means volume 0
            continue
        hdr = rec[:32].view(np.int32)
        n_comp = min(V[u], comp[0][u + 1] - comp[0][u])
        lens = [int(x[0][u + 1] - x[0][u]) for x in (click, conv, rev)]
        cs = comp[1][comp[0][u]:comp[0][u] + n_comp].numpy()
        rs = rev[1][rev[0][u]:rev[0][u] + lens[2]].numpy()
        narrow = int((cs.size == 0 or cs.min() >= 0) and (rs.size == 0 or (rs.min() >= 0 and rs.max() <= 65535)))
        assert hdr.tolist() == [V[u], n_comp, *lens, narrow, 0, 0]   # flags bit 0 = ADC_PACKED_NARROW
        seen_flags.add(narrow)
        pad = (n_comp + 3) & ~3
        c = rec[32:32 + 4 * pad].view(np.int32)
        assert np.array_equal(c[:n_comp], comp[1][comp[0][u]:comp[0][u] + n_comp].numpy())
        assert np.all(c[n_comp:] == np.iinfo(np.int32).max)  # padding never wins an auction
        p = 32 + 4 * pad
        for (o, vals), n, dt, w in ((click, lens[0], np.float64, 8), (conv, lens[1], np.float64, 8),
                                    (rev, lens[2], np.int32, 4)):
            assert np.array_equal(rec[p:p + w * n].view(dt), vals[o[u]:o[u] + n].numpy())
            p += w * n
        assert rec.size == (p + 15) & ~15 and not rec[p:].any()
    assert seen_flags == {0, 1}


def test_trimmed_cuts_streams_to_what_was_consumed():
    from adcraft_b200.tape import DeviceTape
    rng = np.random.default_rng(4)
    E, K = 3, 5
    V, comp, click, conv, rev = _random_csr(rng, E * K, 30)
    t = DeviceTape(torch.from_numpy(V.astype(np.int32)).view(E, K), comp[0], comp[1], click[0], click[1],
                   conv[0], conv[1], rev[0], rev[1])
    I = torch.from_numpy(rng.integers(0, 10, (E, K)))
    cut = t.trimmed(I, I // 2, I // 4)
    want = np.minimum(I.numpy().reshape(-1), np.diff(click[0].numpy()))
    assert np.array_equal(np.diff(cut.click_off.numpy()), want)
    u = int(np.argmax(want))
    assert np.array_equal(cut.u_click[cut.click_off[u]:cut.click_off[u + 1]].numpy(),
                          click[1][click[0][u]:click[0][u] + want[u]].numpy())
    assert np.array_equal(np.diff(cut.comp_off.numpy()), np.minimum(V, np.diff(comp[0].numpy())))
