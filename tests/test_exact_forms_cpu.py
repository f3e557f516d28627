"""The window sampler replaces IEEE division sequences by cheaper forms that are exact by construction (DESIGN.md §3.2,
"instruction-level rules").  The forms are restated here in exact rational arithmetic / numpy float32 and compared with the
quotients they stand for; the device side of the same claims is `tools/div_check.cu` (r_div_nocheck against `/` on the GPU)
and the GPU parity tests.  csrc/mpp_sweep2.cuh: shape_terms, window_visit phase A; csrc/mpp_device.cuh: value_to_class."""
from fractions import Fraction

import numpy as np

F32 = np.float32


def rn32(fr: Fraction) -> np.float32:
    """Fraction -> nearest float32, ties to even (no double rounding: the neighbours of the float64 guess are compared exactly)."""
    f = F32(float(fr))
    cands = [f, np.nextafter(f, F32(np.inf)), np.nextafter(f, F32(-np.inf))]
    return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(F32(c).view(np.uint32)) & 1))


def fma32(a, b, c) -> np.float32:
    return rn32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def test_division_by_three_is_correctly_rounded():
    """shape_terms: q = s * RN(1/3); result = fma(fma(-3, q, s), RN(1/3), q) must equal RN(s / 3) (float(np.mean([d0, d1, d2])) of
    data_energies.py:43 is an IEEE quotient)."""
    rcp = F32(1.0) / F32(3.0)
    assert float(rcp) == 0.3333333432674408  # the constant written in the kernel
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-3, 3, 6000), rng.standard_normal(3000) * 1e-3, rng.standard_normal(1000) * 50,
                         [0.0, 1.0, -1.0, 3.0, 1e-30, 2.9999998]]).astype(F32)
    for x in xs:
        q0 = F32(x * rcp)
        q1 = fma32(fma32(F32(-3.0), q0, x), rcp, q0)
        assert q1 == rn32(Fraction(float(x)) / 3), float(x)


def test_cell_index_by_multiply_shift():
    """window_visit phase A: lane / ncw for ncw <= 6 and lane + 32 < 72 through (x * ceil(65536 / ncw)) >> 16."""
    table = {1: 65536, 2: 32768, 3: 21846, 4: 16384, 5: 13108, 6: 10923}
    for n, mul in table.items():
        assert mul == -(-65536 // n)
        for x in range(0, 96):
            assert (x * mul) >> 16 == x // n


def _edges(i):
    step = (1.0, 1.0 / 32.0, np.pi / 32.0)[i]
    return [F32(k * step) for k in range(33)]


def _class_device(i, v, guess):
    """value_to_class (mpp_device.cuh): a first guess, then the two fix-up loops."""
    e = _edges(i)
    c = min(max(int(guess), 0), 31)
    while c > 0 and v < e[c]:
        c -= 1
    while c < 31 and v >= e[c + 1]:
        c += 1
    return c


def test_class_guess_by_product_equals_quotient_guess():
    """The product v * (1 / step) and the quotient v / step may differ by one in the last place, i.e. by one class at a bin edge; the
    fix-up loops make the result the unique class with edge[c] <= v < edge[c + 1] from either guess (mappings.py:45-74)."""
    rng = np.random.default_rng(3)
    for i, vmax in enumerate((32.0, 1.0, np.pi)):
        step = F32((1.0, 1.0 / 32.0, np.pi / 32.0)[i])
        inv = F32((1.0, 32.0, 32.0 / np.pi)[i])
        edges = _edges(i)
        vals = list(rng.uniform(0, vmax, 3000).astype(F32)) + edges[:32] + [np.nextafter(e, F32(0)) for e in edges[1:32]]
        for v in vals:
            v = F32(v)
            a = _class_device(i, v, np.floor(v / step))
            b = _class_device(i, v, np.floor(v * inv))
            assert a == b
            assert edges[a] <= v and (a == 31 or v < edges[a + 1])


def test_sparse_sum_equals_butterfly_sum():
    """warp_sum_sparse: with at most two non-zero lanes the 5-step xor butterfly returns exactly their float32 sum."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        v = np.zeros(32, F32)
        k = rng.integers(0, 3)
        idx = rng.choice(32, size=k, replace=False)
        v[idx] = (rng.standard_normal(k) * 10 ** rng.uniform(-6, 2)).astype(F32)
        b = v.copy()
        for o in (16, 8, 4, 2, 1):
            b = (b + b[np.arange(32) ^ o]).astype(F32)
        nz = v[v != 0]
        want = F32(0) if len(nz) == 0 else (nz[0] if len(nz) == 1 else F32(nz[0] + nz[1]))
        assert np.all(b == want)
