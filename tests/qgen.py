"""TEST INFRASTRUCTURE: vectorised generators of binary128 test vectors (numpy, bit level)."""
import numpy as np

from qblas_b200 import quad

BIAS = 16383


def mk(sign, e, mh, ml):
    out = np.zeros(np.shape(e) + (2,), dtype=np.uint64)
    out[..., 0] = ml
    out[..., 1] = (np.asarray(sign, dtype=np.uint64) << np.uint64(63)) | (np.asarray(e, dtype=np.uint64) << np.uint64(48)) | (
        np.asarray(mh, dtype=np.uint64) & np.uint64((1 << 48) - 1))
    return out


def mantissas(rng, n):
    """mixture: mostly full random, some double-like / sparse / all-ones / trailing zeros"""
    mh = rng.integers(0, 1 << 48, size=n, dtype=np.uint64)
    ml = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    kind = rng.integers(0, 12, size=n)
    ml = np.where(kind == 5, ml & np.uint64(0xF000000000000000), ml)
    ml = np.where(kind == 6, np.uint64(0), ml)
    mh = np.where(kind == 6, mh & np.uint64(0xFFFFFF000000), mh)
    ml = np.where(kind == 7, np.uint64(0xFFFFFFFFFFFFFFFF), ml)
    mh = np.where(kind == 8, np.uint64(0), mh)
    ml = np.where(kind == 8, np.uint64(0), ml)
    ml = np.where(kind == 9, ml & np.uint64(1), ml)
    mh = np.where(kind == 9, np.uint64(0), mh)
    ml = np.where(kind == 10, ml & ~np.uint64(0xFFFFFFFFFF), ml)
    return mh, ml


def exp_of(q):
    return ((q[..., 1] >> np.uint64(48)) & np.uint64(0x7FFF)).astype(np.int64)


def triples(rng, n, regime):
    """(a, b, c) arrays of n quads for fma(a, b, c); regimes mirror tests/host/q128_host_test.cpp"""
    sa, sb, sc = (rng.integers(0, 2, size=n, dtype=np.uint64) for _ in range(3))
    amh, aml = mantissas(rng, n)
    bmh, bml = mantissas(rng, n)
    cmh, cml = mantissas(rng, n)
    r = rng.integers(0, 1 << 40, size=n)
    clip = lambda e: np.clip(e, 0, 0x7FFE)
    if regime == "any":
        ea, eb, ec = (1 + rng.integers(0, 0x7FFE, size=n) for _ in range(3))
    elif regime == "similar":
        ea, eb, ec = (BIAS - 40 + rng.integers(0, 80, size=n) for _ in range(3))
    elif regime == "cancel":
        ea = BIAS - 20 + rng.integers(0, 40, size=n)
        eb = BIAS - 20 + rng.integers(0, 40, size=n)
        ec = ea + eb - BIAS + rng.integers(-2, 3, size=n)
        sc = (sa ^ sb) ^ np.uint64(1)
    elif regime == "gaps":
        ea = BIAS - 10 + rng.integers(0, 20, size=n)
        eb = BIAS - 10 + rng.integers(0, 20, size=n)
        ec = ea + eb - BIAS + rng.integers(-260, 140, size=n)
    elif regime == "tiny":
        ea = np.where(r % 3 == 0, 0, rng.integers(0, 120, size=n))
        eb = np.where((r >> 3) % 3 == 0, 0, rng.integers(0, 120, size=n))
        ec = np.where((r >> 6) % 3 == 0, 0, rng.integers(0, 120, size=n))
    elif regime == "subres":
        ea = 1 + rng.integers(0, BIAS, size=n)
        eb = clip(BIAS - ea - 60 + rng.integers(0, 200, size=n))
        ec = np.where(r % 4 == 0, 0, rng.integers(0, 140, size=n))
    elif regime == "overflow":
        ea, eb, ec = (0x7FFE - rng.integers(0, 60, size=n) for _ in range(3))
    elif regime == "nearovf":
        ea = BIAS + rng.integers(0, BIAS, size=n)
        eb = np.clip(0x7FFE + BIAS - ea - rng.integers(0, 8, size=n) + 2, 1, 0x7FFE)
        ec = 0x7FFE - rng.integers(0, 130, size=n)
    elif regime == "specials":
        def sp(s, mh, ml):
            k = rng.integers(0, 12, size=n)
            e = BIAS - 3 + rng.integers(0, 6, size=n)
            e = np.where(k == 0, 0, e); mh = np.where(k == 0, 0, mh); ml = np.where(k == 0, 0, ml)       # zero
            e = np.where(k == 1, 0x7FFF, e); mh = np.where(k == 1, 0, mh); ml = np.where(k == 1, 0, ml)  # inf
            e = np.where(k == 2, 0x7FFF, e); mh = np.where(k == 2, 1 << 47, mh)                          # nan
            e = np.where(k == 3, 0, e); mh = np.where(k == 3, 0, mh); ml = np.where(k == 3, 1, ml)       # min subnormal
            e = np.where(k == 4, 0, e)                                                                   # subnormal
            e = np.where(k == 5, 0x7FFE, e)                                                              # huge
            e = np.where(k == 6, 1, e)                                                                   # min normal range
            return e, mh.astype(np.uint64), ml.astype(np.uint64)
        ea, amh, aml = sp(sa, amh, aml)
        eb, bmh, bml = sp(sb, bmh, bml)
        ec, cmh, cml = sp(sc, cmh, cml)
    else:
        raise ValueError(regime)
    a = mk(sa, clip(ea), amh, aml)
    b = mk(sb, clip(eb), bmh, bml)
    c = mk(sc, clip(ec), cmh, cml)
    return a, b, c


REGIMES = ["any", "similar", "cancel", "gaps", "tiny", "subres", "overflow", "nearovf", "specials"]


def matrix(rng, rows, cols, kind="D113", ld=None):
    """rows x cols quads stored with leading dimension ld >= cols (row-walk storage)."""
    ld = cols if ld is None else ld
    buf = quad.random_quads(rng, (rows, ld), "D113" if kind == "pad" else kind)
    return np.ascontiguousarray(buf.reshape(rows * ld, 2))


def reference_benchmark_doubles(count, seed=42):
    """The value stream of the reference's benchmark programs (benchmarks/benchmark.cpp:8-29: std::mt19937(seed) +
    std::uniform_real_distribution<double>(-1, 1), i.e. libstdc++'s generate_canonical<double, 53>: two 32-bit draws, low word first),
    reproduced with numpy's MT19937 (same init_genrand seeding).  Checked against a g++ build of the same three lines."""
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2 ** 32, size=2 * count, dtype=np.uint64).astype(np.float64)
    s = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    s = np.where(s >= 1.0, np.nextafter(1.0, 0.0), s)
    return s * 2.0 - 1.0


def structured_gemm_case(rng, m=70, n=45, k=300):
    """Row-major A (m x k), B (k x n), C (m x n) with structure instead of noise: a shifted upper-triangular A whose rows start with +0 / -0
    (zero products into a zero accumulator), a +0 row and a -0 row, a zero column and -0 rows of B, products that cancel exactly inside a
    panel, specials (Inf, NaN, a subnormal, Inf x 0), negative entries of C.  Shared by the CPU test that pins the oracle against the
    reference on these inputs and the GPU test that runs both reference-order qgemm kernels on them."""
    assert m >= 65 and n >= 31 and k >= 254
    A = matrix(rng, m, k, "Dexp"); B = matrix(rng, k, n, "D113"); C0 = matrix(rng, m, n)
    Am = A.reshape(m, k, 2); Bm = B.reshape(k, n, 2); Cm = C0.reshape(m, n, 2)
    negzero = np.array([0, 0x8000 << 48], dtype=np.uint64)
    for i in range(m):
        Am[i, :min(k, 3 * i)] = 0
        if i % 3 == 1:
            Am[i, :min(k, 3 * i):2] = negzero
    Am[17] = 0; Am[18] = negzero
    Bm[:, 7] = 0; Bm[5::11, :] = negzero
    Am[40, 130:140] = Am[41, 130:140]; Bm[130:135, 3] = Bm[135:140, 3]
    Am[40, 135:140, 1] ^= np.uint64(1 << 63)
    Am[41, 131] = Am[41, 130]; Bm[131, 9] = Bm[130, 9]; Bm[131, 9, 1] ^= np.uint64(1 << 63)
    sa, sb, _ = triples(rng, 16, "specials")
    Am[60:64, 200:204] = sa.reshape(4, 4, 2); Bm[250:254, 20:24] = sb.reshape(4, 4, 2)
    Am[62, 210] = (0, 0x7FFF << 48); Bm[211, 30] = (1, 0x7FFF << 48); Am[63, 220] = (12345, 1 << 40)
    Am[64, 230] = (0, 0x7FFF << 48); Bm[230, :] = 0
    Cm[0, :5, 1] |= np.uint64(1 << 63)
    return A, B, C0
