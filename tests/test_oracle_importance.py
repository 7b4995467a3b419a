"""CPU checks of oracle/importance.py (pcg32 known answers, sampling / merge invariants)."""
import numpy as np

from oracle import importance as oi
from oracle.importance import Pcg32


def _seeded(initstate, initseq):
    # pcg32.h:46-53 seed(): state = 0, inc = (seq << 1) | 1, step, state += initstate, step
    g = Pcg32(0, ((initseq << 1) | 1) & oi.M64)
    g.next_uint()
    g.state = (g.state + initstate) & oi.M64
    g.next_uint()
    return g


def test_pcg32_known_answers():
    # the published demo vector of the PCG reference implementation (pcg32 seeded with 42, stream 54)
    g = _seeded(42, 54)
    assert [g.next_uint() for _ in range(6)] == [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    # the reference's default generator (pcg32.h:32-34) is the generator seeded with the library's default constants
    assert Pcg32().state == oi.PCG_DEFAULT_STATE and Pcg32().inc == oi.PCG_DEFAULT_INC


def test_pcg32_advance_equals_stepping():
    a, b = Pcg32(), Pcg32()
    for _ in range(1000):
        a.next_uint()
    b.advance(1000)
    assert a.state == b.state
    f = [Pcg32().next_float() for _ in range(1)]
    assert 0.0 <= float(f[0]) < 1.0


def _packet(n_rays, rng, max_per_ray=12, empty_every=7):
    counts = rng.integers(2, max_per_ray, n_rays)
    counts[::empty_every] = 0
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    se = np.stack([start, start + counts], 1).astype(np.int32)
    se[counts == 0] = -1
    S = int(counts.sum())
    z = np.zeros((S, 1), np.float32)
    w = np.zeros((S, 1), np.float32)
    for r in range(n_rays):
        if counts[r]:
            z[se[r, 0]:se[r, 1], 0] = np.sort(rng.uniform(0.5, 4.0, counts[r])).astype(np.float32)
            ww = rng.uniform(0, 1, counts[r]).astype(np.float32)
            w[se[r, 0]:se[r, 1], 0] = ww / ww.sum()
    return se, z, w


def test_importance_sample_invariants():
    from oracle import compositing as oc

    rng = np.random.default_rng(3)
    se, z, w = _packet(64, rng)
    cdf = oc.packed_compute_cdf(se, w)
    o = rng.normal(size=(64, 3)).astype(np.float32)
    d = rng.normal(size=(64, 3)).astype(np.float32)
    for jitter in (False, True):
        out = oi.importance_sample(o, d, se, z, cdf, 8, jitter)
        for r in range(64):
            s, e = out["ray_start_end_idx"][r]
            if se[r, 1] - se[r, 0] == 0:
                assert (s, e) == (-1, -1)
                continue
            assert (s, e) == (r * 8, r * 8 + 8)
            zi = out["samples_z"][s:e, 0]
            zr = z[se[r, 0]:se[r, 1], 0]
            assert np.all(np.diff(zi) >= 0) and zi.min() >= zr.min() - 1e-6 and zi.max() <= zr.max() + 1e-6
            assert np.allclose(out["samples_3d"][s:e], o[r] + zi[:, None] * d[r], rtol=1e-6, atol=1e-6)


def test_combine_invariants():
    rng = np.random.default_rng(5)
    se1, z1, _ = _packet(50, rng)
    se2, z2, _ = _packet(50, rng, max_per_ray=6, empty_every=5)
    mk = lambda z, off: (np.arange(off, off + len(z), dtype=np.int32).reshape(-1, 1), rng.normal(size=(len(z), 3)).astype(np.float32),
                         rng.normal(size=(len(z), 3)).astype(np.float32), rng.normal(size=(len(z), 2)).astype(np.float32))
    i1, p1, d1, v1 = mk(z1, 0)
    i2, p2, d2, v2 = mk(z2, 10000)
    for md in (0.0, 0.05):
        out = oi.combine_ray_samples_packets(se1, i1, p1, d1, z1, v1, se2, i2, p2, d2, z2, v2, md)
        for r in range(50):
            a = z1[se1[r, 0]:se1[r, 1], 0] if se1[r, 1] > se1[r, 0] else np.zeros(0, np.float32)
            b = z2[se2[r, 0]:se2[r, 1], 0] if se2[r, 1] > se2[r, 0] else np.zeros(0, np.float32)
            s, e = out["ray_start_end_idx"][r]
            if len(a) + len(b) == 0:
                assert (s, e) == (-1, -1)
                continue
            zc = out["samples_z"][s:e, 0]
            assert np.all(np.diff(zc) >= md - 1e-7)
            if md == 0.0:
                assert np.array_equal(zc, np.sort(np.concatenate([a, b])))
            # every kept sample carries its source row's payload
            for k in range(s, e):
                idx = int(out["samples_idx"][k, 0])
                src_p = p1[idx] if idx < 10000 else p2[idx - 10000]
                assert np.array_equal(out["samples_3d"][k], src_p)
