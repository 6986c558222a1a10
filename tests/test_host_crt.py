"""CPU checks of the residue arithmetic of the tensor-path qgemm (qblas_b200/csrc/qb_crt.cuh, host/device
dual source, built with g++ through tests/host/crt_host.cpp) against Python integers: int8 residues of
signed multi-word integers, reduction of int32 accumulators, and the grouped Chinese-remainder
reconstruction of the exact inner product."""
import ctypes as C
import os
import random
import subprocess
from math import gcd, log2

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("crt") / "libcrt_host.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=gnu++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "host", "crt_host.cpp")], check=True)
    L = C.CDLL(str(so))
    L.crt_plan_bits.restype = C.c_double
    L.crt_residues.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.crt_fold.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.crt_set_form.argtypes = [C.c_int]
    return L


def _moduli(lib):
    return [lib.crt_modulus(i) for i in range(lib.crt_num_moduli())]


def test_moduli_pairwise_coprime_and_int8(lib):
    m = _moduli(lib)
    assert m == sorted(m, reverse=True) and max(m) <= 256
    for i in range(len(m)):
        for j in range(i):
            assert gcd(m[i], m[j]) == 1
    for g in range(0, len(m), 4):
        prod = 1
        for p in m[g:g + 4]:
            prod *= p
        assert prod < 2 ** 32
    # capacity: full-mantissa operands (W ~ 142) at k = 2^15 and beyond
    tot = sum(log2(p) for p in m)
    assert tot > 142 * 2 + 15 + 1
    assert abs(lib.crt_plan_bits(len(m)) - tot) < 1e-6
    assert lib.crt_moduli_for_bits(292) == 41 and lib.crt_moduli_for_bits(400) == 0


def _words(x):
    return [(x >> (32 * j)) & 0xffffffff for j in range(6)]


def _residues(lib, xs, nw, N):
    words = np.array([_words(abs(x)) for x in xs], dtype=np.uint32)
    sign = np.array([1 if x < 0 else 0 for x in xs], dtype=np.uint32)
    out = np.zeros((len(xs), N), dtype=np.int8)
    lib.crt_residues(len(xs), words.ctypes.data, sign.ctypes.data, nw, N, out.ctypes.data)
    return out


@pytest.mark.parametrize("W", [1, 8, 31, 32, 33, 54, 113, 139, 160, 192])
def test_residues_are_symmetric_int8(lib, W):
    rnd = random.Random(W)
    m = _moduli(lib)
    xs = [0, 1, -1, 2 ** W - 1, -(2 ** W - 1)] + [rnd.randrange(-(2 ** W) + 1, 2 ** W) for _ in range(300)]
    nw = (W + 31) // 32
    out = _residues(lib, xs, nw, len(m))
    for x, row in zip(xs, out):
        for p, r in zip(m, row):
            r = int(r)
            assert (r - x) % p == 0
            assert -128 <= r <= 127 and (-(p - 1) // 2 <= r <= (p - 1) // 2 if p % 2 else -128 <= r <= 127)


@pytest.mark.parametrize("WA,WB,k", [(139, 139, 64), (142, 137, 300), (54, 54, 128), (113, 113, 17), (192, 120, 8), (7, 5, 1000), (160, 160, 50)])
def test_inner_product_reconstruction_exact(lib, WA, WB, k):
    rnd = random.Random(WA * 1000 + WB + k)
    need = WA + WB + (k - 1).bit_length() + 1
    N = lib.crt_moduli_for_bits(need)
    assert N > 0
    m = _moduli(lib)[:N]
    cases = 24
    acc = np.zeros((cases, N), dtype=np.int32)
    exact = []
    for c in range(cases):
        if c == 0:      # extreme: every product at the bound, same sign
            a = [2 ** WA - 1] * k; b = [-(2 ** WB - 1)] * k
        elif c == 1:    # exact cancellation
            a = [rnd.randrange(-(2 ** WA) + 1, 2 ** WA) for _ in range(k // 2)]; a = a + a + [0] * (k - 2 * (k // 2))
            b = [rnd.randrange(-(2 ** WB) + 1, 2 ** WB) for _ in range(k // 2)]; b = b + [-v for v in b] + [0] * (k - 2 * (k // 2))
        else:
            a = [rnd.randrange(-(2 ** WA) + 1, 2 ** WA) for _ in range(k)]
            b = [rnd.randrange(-(2 ** WB) + 1, 2 ** WB) for _ in range(k)]
        ra = _residues(lib, a, (WA + 31) // 32, N).astype(np.int64)
        rb = _residues(lib, b, (WB + 31) // 32, N).astype(np.int64)
        acc[c] = (ra * rb).sum(axis=0)       # what one int8 GEMM per modulus accumulates (|.| <= k 2^14)
        exact.append(sum(x * y for x, y in zip(a, b)))
    mag = np.zeros((cases, 14), dtype=np.uint32); neg = np.zeros(cases, dtype=np.uint32)
    lib.crt_fold(cases, acc.ctypes.data, N, mag.ctypes.data, neg.ctypes.data)
    for c in range(cases):
        v = sum(int(mag[c, l]) << (32 * l) for l in range(14))
        v = -v if neg[c] else v
        assert v == exact[c], (c, N)
    assert exact[1] == 0 and neg[1] == 0


def test_accumulator_reduction_range(lib):
    """int32 accumulators up to +-2^30 (k = 65536 products of +-128 * +-128) reduce exactly."""
    N = 8
    m = _moduli(lib)[:N]
    vals = [0, 1, -1, 2 ** 30, -(2 ** 30), 2 ** 30 - 1, -(2 ** 30) + 1, 12345678, -87654321]
    # reconstruction of small integers straight from their accumulators
    acc = np.array([[v] * N for v in vals], dtype=np.int32)
    mag = np.zeros((len(vals), 14), dtype=np.uint32); neg = np.zeros(len(vals), dtype=np.uint32)
    lib.crt_fold(len(vals), acc.ctypes.data, N, mag.ctypes.data, neg.ctypes.data)
    for c, v in enumerate(vals):
        got = sum(int(mag[c, l]) << (32 * l) for l in range(14))
        assert (-got if neg[c] else got) == v


@pytest.mark.parametrize("form", [0, 1], ids=["reference-form", "kernel-form"])
def test_reconstruction_every_moduli_count(lib, form):
    """Every N = 1..49 (group counts 1..13, full and partial last groups): integers spread over (-P/2, P/2), including the
    extremes, are recovered exactly from their residues - by the reference form and by the kernel's carry-chain form
    (crt::reconstruct_dev, compiled here with the PTX carry flag emulated)."""
    rnd = random.Random(1234 + form)
    m_all = _moduli(lib)
    lib.crt_set_form(form)
    try:
        for N in range(1, len(m_all) + 1):
            m = m_all[:N]
            P = 1
            for p in m:
                P *= p
            half = (P - 1) // 2            # |I| < P/2 strictly (P is even: 256 is always among the moduli)
            vals = [0, 1, -1, half, -half, half - 1, 1 - half] + [rnd.randrange(-half, half + 1) for _ in range(40)] + \
                   [rnd.randrange(-(2 ** b), 2 ** b) for b in range(1, max(2, half.bit_length() - 1), 7)]
            # accumulators: any int32 congruent to the value (the tensor kernel hands over sums of products, not residues)
            acc = np.zeros((len(vals), N), dtype=np.int32)
            for c, v in enumerate(vals):
                for i, p in enumerate(m):
                    r = v % p
                    kmax = (2 ** 30 - r) // p
                    acc[c, i] = r + p * rnd.randrange(-kmax, kmax + 1) if c % 2 else r
            mag = np.zeros((len(vals), 14), dtype=np.uint32); neg = np.zeros(len(vals), dtype=np.uint32)
            lib.crt_fold(len(vals), acc.ctypes.data, N, mag.ctypes.data, neg.ctypes.data)
            for c, v in enumerate(vals):
                got = sum(int(mag[c, l]) << (32 * l) for l in range(14))
                assert (-got if neg[c] else got) == v, (N, c, v)
    finally:
        lib.crt_set_form(0)


def test_inner_products_through_kernel_form(lib):
    """The exact-inner-product check of above, with the kernel's form of the reconstruction."""
    lib.crt_set_form(1)
    try:
        for WA, WB, k in [(139, 139, 64), (54, 54, 128), (24, 24, 40), (80, 11, 300), (7, 5, 1000)]:
            test_inner_product_reconstruction_exact(lib, WA, WB, k)
    finally:
        lib.crt_set_form(0)


@pytest.mark.parametrize("nl", [2, 3, 5, 9, 10, 11, 12, 14])
def test_one_rounding_of_the_wide_integer(lib, nl):
    """crt::limbs_to_q: +-|I| * 2^Eb -> binary128 with ONE round-to-nearest-even, for magnitudes of every length (short ones,
    exactly representable ones, ties with even / odd mantissas, sticky bits far below the window, all-ones carries), in the normal
    range, in gradual underflow and at overflow - against exact rational arithmetic."""
    from fractions import Fraction
    from qblas_b200 import quad
    lib.crt_round.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rnd = random.Random(nl)
    cases = []
    maxbits = 32 * nl
    for _ in range(400):
        bits = rnd.randrange(1, maxbits + 1)
        kind = rnd.randrange(6)
        if kind == 0:
            v = rnd.getrandbits(bits) | (1 << (bits - 1))
        elif kind == 1:   # 113 significant bits then a tie pattern: ...1000 / ...0111 / ...1001
            top = rnd.getrandbits(113) | (1 << 112)
            low = max(0, bits - 113)
            tail = [1 << (low - 1), (1 << (low - 1)) - 1, (1 << (low - 1)) + 1][rnd.randrange(3)] if low >= 2 else 0
            v = (top << low) | tail
        elif kind == 2:   # all ones: rounding carries into the next binade
            v = (1 << bits) - 1
        elif kind == 3:   # a power of two, or one with a single far-away sticky bit
            v = (1 << (bits - 1)) | (1 if rnd.randrange(2) else 0)
        elif kind == 4:   # short value
            v = rnd.getrandbits(min(bits, 40)) or 1
        else:             # tie exactly at half an ulp with a random (even or odd) mantissa
            top = rnd.getrandbits(113) | (1 << 112)
            low = max(1, bits - 113)
            v = (top << low) | (1 << (low - 1))
        v &= (1 << maxbits) - 1
        v = v or 1
        reg = rnd.randrange(4)
        Eb = [rnd.randrange(-400, 400), rnd.randrange(-16494 - 120, -16494 + 40) - v.bit_length() + 113,
              16383 - v.bit_length() + rnd.randrange(-3, 3), rnd.randrange(-16600, -16300)][reg]
        cases.append((v, rnd.randrange(2), Eb))
    cases.append((0, 0, 0)); cases.append((0, 1, -50))
    mag = np.zeros((len(cases), 14), dtype=np.uint32)
    for c, (v, _, _) in enumerate(cases):
        for l in range(nl):
            mag[c, l] = (v >> (32 * l)) & 0xffffffff
    neg = np.array([c[1] for c in cases], dtype=np.uint32); Eb = np.array([c[2] for c in cases], dtype=np.int32)
    out = np.zeros((len(cases), 2), dtype=np.uint64)
    lib.crt_round(len(cases), nl, mag.ctypes.data, neg.ctypes.data, Eb.ctypes.data, out.ctypes.data)
    for c, (v, s, e) in enumerate(cases):
        if v == 0:
            assert (int(out[c, 1]), int(out[c, 0])) == (0, 0)        # +0, whatever the sign flag says
            continue
        hi, lo = quad.from_fraction(Fraction(-v if s else v) * Fraction(2) ** e)
        if hi & 0x7fffffffffffffff == 0 and lo == 0:
            hi |= s << 63                                            # from_fraction drops the sign of an underflow to zero
        assert (int(out[c, 1]), int(out[c, 0])) == (hi, lo), (nl, c, hex(v), s, e)


@pytest.mark.parametrize("kind,m,n,k", [("D113", 5, 4, 37), ("D53", 4, 6, 64), ("Dint", 3, 5, 20), ("Dexp8", 4, 3, 33), ("D113", 2, 2, 300),
                                       ("Dzero", 3, 3, 10), ("Dfloat", 4, 4, 50)])
def test_whole_scheme_on_the_cpu_is_exactly_rounded(lib, oracle, kind, m, n, k):
    """The residue-scheme qgemm end to end on the CPU, stitched from the library's own dual-source arithmetic (row / column scan,
    element -> integer words, int8 residues, int32 accumulation per modulus, reduction, carry-chain reconstruction, the one
    rounding, the reference epilogue): bit for bit the exact inner products rounded once (tests/exact_ref.py), then
    C = fma(alpha, s, mul(beta, C)) by the oracle."""
    from exact_ref import exact_matmul_rounded
    from qblas_b200 import quad
    rng = np.random.default_rng(m * 100 + n * 10 + k)

    def mk(r, c):
        if kind == "Dint":
            return quad.from_double(rng.integers(-9, 10, size=(r, c)).astype(np.float64)).reshape(r * c, 2)
        if kind == "Dexp8":
            return np.ascontiguousarray(quad.random_quads(rng, (r, c), "D113", emin=-8, emax=8).reshape(r * c, 2))
        if kind == "Dzero":
            return quad.from_double(np.zeros((r, c))).reshape(r * c, 2)
        if kind == "Dfloat":   # float32 values: 24-bit spans -> few moduli, two reconstruction groups
            return quad.from_double(rng.standard_normal((r, c)).astype(np.float32).astype(np.float64)).reshape(r * c, 2)
        return np.ascontiguousarray(quad.random_quads(rng, (r, c), kind).reshape(r * c, 2))
    A = mk(m, k); B = mk(k, n) if kind != "Dzero" else np.ascontiguousarray(quad.random_quads(rng, (k, n), "D113").reshape(k * n, 2))
    C0 = np.ascontiguousarray(quad.random_quads(rng, (m, n), "D113").reshape(m * n, 2))
    alpha, beta = quad.random_quads(rng, 2)
    got = C0.copy()
    info = np.zeros(4, dtype=np.int32)
    rej = np.zeros(m * n, dtype=np.uint8)
    lib.crt_set_form(1)
    lib.crt_gemm_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    al = np.ascontiguousarray(alpha, dtype=np.uint64); be = np.ascontiguousarray(beta, dtype=np.uint64)
    rc = lib.crt_gemm_host(m, n, k, A.ctypes.data, B.ctypes.data, got.ctypes.data, al.ctypes.data, be.ctypes.data, 144, info.ctypes.data, rej.ctypes.data)
    assert rc == 0, "scheme declined"
    assert info[3] == 0 and not rej.any()          # windows cover the spans: nothing truncated, nothing rejected
    s = exact_matmul_rounded(A, k, B, n, m, n, k)
    alb = np.broadcast_to(al.reshape(1, 2), s.shape).copy(); beb = np.broadcast_to(be.reshape(1, 2), s.shape).copy()
    want = oracle.fma(alb, s, oracle.mul(beb, C0))
    assert quad.same_bits(got, want).all(), (kind, info.tolist())
    if kind == "D53":
        assert info[1] <= 64 and info[2] <= 64 and info[0] <= 17      # doubles: short spans, few moduli
    if kind == "Dint":
        assert info[0] <= 4


@pytest.mark.parametrize("kind,m,n,k,wcap", [("Dexp40", 6, 5, 64, 144), ("Dexp40", 4, 4, 200, 128), ("Dexp40xD53", 5, 5, 80, 144), ("Dexp70xD53", 5, 5, 80, 144), ("Dexp200", 4, 6, 48, 144),
                                            ("cancel", 4, 4, 64, 144), ("Dexp40", 3, 3, 2, 144)])
def test_capped_windows_meet_the_contract_or_are_rejected(lib, oracle, kind, m, n, k, wcap):
    """Exponent spreads the moduli cannot cover (SURVEY.md §8d cfg3 'Dexp' = D113 x 2^U{-40..40}, and far wider): the planner caps the
    windows, element_words drops the low bits of the small elements, and k_crt_fold's acceptance test (crt::accept_msb) decides per
    element.  Here the same source runs on the CPU: every ACCEPTED element is inside |c^ - c| <= k u (|A||B|)_ij against exact
    long-accumulator arithmetic (oracle/qoracle.c), every other one is reported for the fix-up; engineered cancellations are caught."""
    from qblas_b200 import quad
    rng = np.random.default_rng(m * 1000 + n * 100 + k + wcap)

    def mk(r, c, spread):
        return np.ascontiguousarray(quad.random_quads(rng, (r, c), "D113", emin=-spread, emax=spread).reshape(r * c, 2))
    if kind == "Dexp40":
        A, B = mk(m, k, 40), mk(k, n, 40)
    elif kind in ("Dexp40xD53", "Dexp70xD53"):
        A = mk(m, k, 40 if kind == "Dexp40xD53" else 70); B = np.ascontiguousarray(quad.random_quads(rng, (k, n), "D53").reshape(k * n, 2))
    elif kind == "Dexp200":
        A, B = mk(m, k, 200), mk(k, n, 200)
    else:   # rows whose large terms cancel exactly: what is left comes from the small elements the cap truncates
        A, B = mk(m, k, 40), mk(k, n, 40)
        Am = A.reshape(m, k, 2); Bm = B.reshape(k, n, 2)
        big = quad.from_double(np.array([2.0 ** 60]))[0]
        Am[:, 0] = big; Am[:, 1] = big; Am[:, 1, 1] ^= np.uint64(1 << 63)      # +2^60, -2^60
        Bm[1, :] = Bm[0, :]                                                    # times equal entries: the pair cancels exactly
    C0 = np.ascontiguousarray(quad.random_quads(rng, (m, n), "D113").reshape(m * n, 2))
    got = C0.copy()
    info = np.zeros(4, dtype=np.int32)
    rej = np.zeros(m * n, dtype=np.uint8)
    one = quad.from_double(np.array([1.0])); zero = quad.from_double(np.array([0.0]))
    lib.crt_set_form(1)
    lib.crt_gemm_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rc = lib.crt_gemm_host(m, n, k, A.ctypes.data, B.ctypes.data, got.ctypes.data, one.ctypes.data, zero.ctypes.data, wcap, info.ctypes.data, rej.ctypes.data)
    assert rc == 0
    N, WA, WB, trunc = info.tolist()
    assert WA + WB <= 2 * wcap and max(WA, WB) <= 192, info.tolist()
    idx = np.stack(np.meshgrid(np.arange(m), np.arange(n), indexing="ij"), axis=-1).reshape(-1, 2)
    exact, ratio, klass = oracle.exact_dot_check("R", k, A, k, B, n, idx, got)
    if kind == "Dexp40xD53":                     # the narrow operand leaves its share of the budget to the wide one: still exact
        assert trunc == 0 and WB <= 60 and WA > wcap and not rej.any() and quad.same_bits(got, exact).all(), info.tolist()
        return
    assert trunc != 0, info.tolist()
    if kind == "Dexp70xD53":
        assert trunc == 1 and WA == 192 and WB <= 60, info.tolist()
    acc = rej == 0
    assert (klass == 0).all()
    assert (ratio[acc] <= 1.0).all(), (info.tolist(), ratio[acc].max())
    assert (got[~acc] == C0[~acc]).all()                                       # rejected elements are left to the fix-up
    if kind == "cancel":
        assert rej.sum() >= 1, "the engineered cancellations must be rejected"
    elif k > 1 and wcap >= 144 and kind != "Dexp200":
        # (a 128-bit window leaves ~10 bits of margin over the test, and with +-200 binades the largest element of a row rarely meets
        # the largest of a column, so the sums sit far below the windows' product: many rejections, all legitimate)
        assert rej.mean() < 0.25, rej.mean()
