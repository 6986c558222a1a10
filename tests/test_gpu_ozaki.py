"""GPU tests of the fast-mode tensor-core qgemm (csrc/qb_ozaki.cu): the int8 tcgen05 kernel alone
against integer matmul, and the whole path (scan -> slice -> mma -> fold) against EXACT inner
products rounded once (tests/exact_ref.py) followed by the reference epilogue
C = fma(alpha, s, mul(beta, C)) (/root/reference/include/quadblas/algorithms/level3.hpp:102-109)."""
import numpy as np
import pytest
import torch

import qgen
from exact_ref import exact_matmul_rounded
from gpu_util import dev_random, to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_setting(qb):
    """The tests of this file that compare bit for bit run the digit-diagonal scheme with ALL diagonals
    (qb_set_tensor_keep(0)); the bounded setting is covered by the test_bounded_* tests, which select it themselves.
    The residue scheme (qb_set_tensor_scheme(1), the default) is always exact."""
    old, old_scheme = qb.get_tensor_keep(), qb.get_tensor_scheme()
    qb.set_tensor_keep(0)
    yield
    qb.set_tensor_keep(old); qb.set_tensor_scheme(old_scheme)


SCHEMES = [pytest.param(1, id="residues"), pytest.param(0, id="digits")]


def _diag_ref(pa, pb, m, n):
    SA, SB = pa.shape[0], pb.shape[0]
    a = pa.cpu().numpy().astype(np.int64)
    b = pb.cpu().numpy().astype(np.int64)
    out = np.zeros((SA + SB - 1, m, n), dtype=np.int64)
    for s in range(SA):
        for t in range(SB):
            out[s + t] += a[s, :m] @ b[t, :n].T
    return out


@pytest.mark.parametrize("SA,SB,m,n,Kp", [(1, 1, 128, 256, 128), (1, 1, 128, 256, 512), (3, 2, 200, 300, 384), (2, 5, 130, 513, 1024),
                                           (4, 4, 640, 1024, 256)])
def test_i8_diagonal_gemm_kernel(qb, SA, SB, m, n, Kp):
    g = torch.Generator(device="cuda"); g.manual_seed(SA * 100 + SB)
    pa = torch.randint(-128, 128, (SA, m, Kp), generator=g, device="cuda", dtype=torch.int8)
    pb = torch.randint(-128, 128, (SB, n, Kp), generator=g, device="cuda", dtype=torch.int8)
    Mp, Np = (m + 127) // 128 * 128, (n + 255) // 256 * 256
    D = torch.full((SA + SB - 1, Mp, Np), 77, device="cuda", dtype=torch.int32)
    qb.oz_i8gemm(pa, pb, m, n, D)
    torch.cuda.synchronize()
    ref = _diag_ref(pa, pb, m, n)
    got = D.cpu().numpy()[:, :m, :n].astype(np.int64)
    assert (got == ref).all(), f"mismatch: {np.argwhere(got != ref)[:5]}"
    # rows / columns beyond m, n come from TMA zero fill
    assert (D.cpu().numpy()[:, m:, :] == 0).all() and (D.cpu().numpy()[:, :, n:] == 0).all()


def test_i8_kernel_k_chunks(qb):
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    SA, SB, m, n, Kp = 2, 3, 256, 256, 1024
    pa = torch.randint(-128, 128, (SA, m, Kp), generator=g, device="cuda", dtype=torch.int8)
    pb = torch.randint(-128, 128, (SB, n, Kp), generator=g, device="cuda", dtype=torch.int8)
    D1 = torch.zeros((SA + SB - 1, m, n), device="cuda", dtype=torch.int32)
    D2 = torch.zeros_like(D1)
    qb.oz_i8gemm(pa, pb, m, n, D1, kb_begin=0, nkb=3)
    qb.oz_i8gemm(pa, pb, m, n, D2, kb_begin=3, nkb=5)
    torch.cuda.synchronize()
    ref = _diag_ref(pa, pb, m, n)
    assert ((D1 + D2).cpu().numpy().astype(np.int64) == ref).all()


def _epilogue(oracle, alpha, s, beta, C0):
    al = np.broadcast_to(np.asarray(alpha, dtype=np.uint64).reshape(1, 2), s.shape).copy()
    be = np.broadcast_to(np.asarray(beta, dtype=np.uint64).reshape(1, 2), s.shape).copy()
    return oracle.fma(al, s, oracle.mul(be, np.ascontiguousarray(C0)))


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("m,n,k,kind,layout", [(40, 33, 300, "D113", "R"), (17, 50, 129, "Dexp", "R"), (33, 20, 257, "D53", "C"),
                                               (130, 260, 64, "D113", "R"), (5, 7, 1000, "D113", "C"), (24, 30, 140, "Dexp8", "R"),
                                               (9, 300, 70, "Dint", "C"), (140, 12, 33, "D113", "C")])
def test_fast_gemm_tensor_path_is_exactly_rounded(qb, oracle, m, n, k, kind, layout, scheme):
    qb.set_tensor_scheme(scheme)
    rng = np.random.default_rng(m + n + k)
    ar, ac = (m, k) if layout == "R" else (k, m)
    br, bc = (k, n) if layout == "R" else (n, k)
    cr, cc = (m, n) if layout == "R" else (n, m)
    lda, ldb, ldc = ac + 1, bc + 2, cc + 3
    def mk(r, c, ld):
        if kind == "Dexp":  # +-28 binades: 113 + 56 + 2 bits -> 22 digits (the full +-40 of qgen needs 25 > 24 and is declined);
            # 2 x 171 bits + log2 k is also more than the residue scheme's 49 moduli cover: it hands over to the digit diagonals
            return np.ascontiguousarray(quad.random_quads(rng, (r, ld), "D113", emin=-28, emax=28).reshape(r * ld, 2))
        if kind == "Dexp8":  # +-8 binades: spans of ~130 bits, inside the residue scheme
            return np.ascontiguousarray(quad.random_quads(rng, (r, ld), "D113", emin=-8, emax=8).reshape(r * ld, 2))
        if kind == "Dint":   # small integers (a few moduli, one reconstruction group), with zeros
            return quad.from_double(rng.integers(-9, 10, size=(r, ld)).astype(np.float64)).reshape(r * ld, 2)
        return qgen.matrix(rng, r, c, kind, ld)
    A = mk(ar, ac, lda); B = mk(br, bc, ldb); C0 = mk(cr, cc, ldc)
    alpha, beta = quad.random_quads(rng, 2)
    s = exact_matmul_rounded(A, lda, B, ldb, m, n, k, layout)          # (m*n, 2) in (i, j) order
    idx = np.array([[(i * ldc + j) if layout == "R" else (j * ldc + i) for j in range(n)] for i in range(m)]).reshape(-1)
    want = _epilogue(oracle, alpha, s, beta, C0[idx])
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
    try:
        dC = to_dev(C0)
        qb.gemm(layout, m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    got = to_host(dC)
    assert st["pairs"] > 0, "tensor path declined"
    assert st["scheme"] == ("residues" if scheme == 1 and kind != "Dexp" else "digits"), st
    assert quad.same_bits(got[idx], want).all(), f"{(~quad.same_bits(got[idx], want)).sum()} mismatches, plan {st}"
    # untouched padding
    mask = np.ones(C0.shape[0], dtype=bool); mask[idx] = False
    assert (got[mask] == C0[mask]).all()


@pytest.mark.parametrize("scheme", SCHEMES)
def test_fast_gemm_cancellation_pattern_is_exact(qb, oracle, scheme):
    """README:141-158 / test_quadblas.cpp:715-739: rows (1e20, 1, -1e20, 0...) times ones -> exactly 1."""
    qb.set_tensor_scheme(scheme)
    m, n, k = 128, 128, 256
    A = np.zeros((m, k)); A[:, 0] = 1e20; A[:, 1] = 1.0; A[:, 2] = -1e20
    Aq = quad.from_double(A).reshape(-1, 2); Bq = quad.from_double(np.ones((k, n))).reshape(-1, 2)
    C0 = quad.from_double(np.zeros((m, n))).reshape(-1, 2)
    qb.set_mode(qb.MODE_FAST)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, 1.0, to_dev(Aq), k, to_dev(Bq), n, 0.0, dC, n)
        torch.cuda.synchronize()
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    one = quad.from_double(np.ones(m * n))
    assert quad.same_bits(to_host(dC), one).all()


@pytest.mark.parametrize("scheme", SCHEMES)
def test_fast_gemm_declines_specials_and_wide_spans(qb, oracle, scheme):
    """Inf/NaN or a row spanning more than 24 digits: the planner declines, the integer kernel runs (fast mode = single chain)."""
    qb.set_tensor_scheme(scheme)
    rng = np.random.default_rng(3)
    m, n, k = 130, 140, 260
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    A[5] = quad.from_double(np.array([np.inf]))[0]
    A[700] = quad.from_double(np.array([1e-300]))[0] ; A[701] = quad.from_double(np.array([1e300]))[0]
    Co = C0.copy()
    oracle.gemm("R", m, n, k, 1.0, A, k, B, n, 1.0, Co, n, kc=k)       # single chain = fast-mode integer kernel
    qb.set_mode(qb.MODE_FAST)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, 1.0, to_dev(A), k, to_dev(B), n, 1.0, dC, n)
        torch.cuda.synchronize()
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    assert quad.same_bits(to_host(dC), Co).all()


@pytest.mark.parametrize("scheme", SCHEMES)
def test_fast_gemm_large_matches_sampled_exact(qb, oracle, scheme):
    """1024 x 768 x 2304 (two K chunks at 18 slices would need k > 7281; force chunks via D113 + k): sampled entries vs exact."""
    qb.set_tensor_scheme(scheme)
    m, n, k = 1024, 768, 2304
    A = dev_random((m * k,), "D113", seed=1); B = dev_random((k * n,), "D113", seed=2); C = dev_random((m * n,), "D113", seed=3)
    C0 = to_host(C).copy()
    qb.set_mode(qb.MODE_FAST)
    try:
        qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    assert st["pairs"] > 0 and st["scheme"] == ("residues" if scheme else "digits"), st
    Ah, Bh, got = to_host(A), to_host(B), to_host(C)
    rng = np.random.default_rng(0)
    for _ in range(12):
        i, j = int(rng.integers(m)), int(rng.integers(n))
        s = exact_matmul_rounded(Ah[i * k:(i + 1) * k], k, np.ascontiguousarray(Bh[j::n][:k]), 1, 1, 1, k)
        want = _epilogue(oracle, quad.from_double(np.array([1.0]))[0], s, quad.from_double(np.array([0.0]))[0], C0[i * n + j:i * n + j + 1])
        assert quad.same_bits(got[i * n + j:i * n + j + 1], want).all(), (i, j, st)


# ------------------------------------------------------------------ bounded setting (qb_set_tensor_keep(d), default d = 17)
def _contract_check(qb, oracle, m, n, k, A, B, got, C0=None, beta0=True):
    """|c^ - c| <= gamma_k (|A||B|)_ij for every entry, c = exact inner product (alpha = 1, beta = 0)."""
    from fractions import Fraction
    s = exact_matmul_rounded(A, k, B, n, m, n, k, "R")                 # exact, rounded once: within u|c| of c
    idx = np.stack(np.meshgrid(np.arange(m), np.arange(n), indexing="ij"), axis=-1).reshape(-1, 2)
    ab = oracle.absdot_sample("R", k, A, k, B, n, idx)
    u = Fraction(1, 2 ** 113); gam = k * u / (1 - k * u)
    f = lambda v: quad.to_fraction(int(v[1]), int(v[0]))
    worst = Fraction(0)
    for q in range(m * n):
        err = abs(f(got[q]) - f(s[q]))
        bound = gam * f(ab[q])
        assert err <= bound + u * abs(f(s[q])), (q, float(err), float(bound))
        if bound: worst = max(worst, err / bound)
    return float(worst)


@pytest.mark.parametrize("m,n,k,kind", [(130, 257, 300, "D113"), (128, 256, 1024, "D113"), (140, 260, 520, "Dexp")])
def test_bounded_tensor_path_meets_contract(qb, oracle, m, n, k, kind):
    rng = np.random.default_rng(m * 7 + k)
    mk = (lambda r, c: np.ascontiguousarray(quad.random_quads(rng, (r, c), "D113", emin=-28, emax=28).reshape(r * c, 2))) if kind == "Dexp" \
        else (lambda r, c: qgen.matrix(rng, r, c, kind, c))
    A = mk(m, k); B = mk(k, n); C0 = qgen.matrix(rng, m, n, "D113", n)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_keep(17); qb.set_tensor_scheme(0)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, 1.0, to_dev(A), k, to_dev(B), n, 0.0, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
        dC2 = to_dev(C0)
        qb.gemm("R", m, n, k, 1.0, to_dev(A), k, to_dev(B), n, 0.0, dC2, n)
        torch.cuda.synchronize()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["pairs"] > 0 and st["keep"] == 17 and st["keep"] < st["ndiag"], st
    want_pairs = sum(min(d, st["SA"] - 1) - max(0, d - (st["SB"] - 1)) + 1 for d in range(17))
    assert st["pairs"] == want_pairs < st["SA"] * st["SB"], st        # e.g. 153 digit-plane products instead of 18 * 18
    got = to_host(dC)
    assert quad.same_bits(got, to_host(dC2)).all()                    # deterministic
    worst = _contract_check(qb, oracle, m, n, k, A, B, got)
    assert worst < 0.5                                                # typically ~1/k: far inside the bound


def test_bounded_fixup_on_cancelling_entries(qb, oracle):
    """Columns of B built so that some inner products cancel to ~2^-60 of their terms: those entries fail the
    |J| >= 2^125 check, are left to k_oz_fixup and must still meet the contract."""
    rng = np.random.default_rng(77)
    m, n, k = 128, 256, 512
    A = qgen.matrix(rng, m, k, "D113", k)
    Bm = quad.random_quads(rng, (k, n), "D113")
    # column j (odd) = -(column j-1) in the first half of k, + the same in the second half, applied to rows of A
    # that repeat their first half: A[i, k/2 + l] = A[i, l] for i < 4  ->  c[i, j] cancels to ~2^-107 for those (i, j odd)
    Am = A.reshape(m, k, 2).copy()
    Am[:4, k // 2:] = Am[:4, :k // 2]
    Bm[k // 2:, 1::2] = Bm[:k // 2, 1::2] ^ np.array([0, 1 << 63], dtype=np.uint64)   # negated copy
    Bm[k // 2:, 1::2, 0] ^= np.uint64(1) << np.uint64(5)                               # ... up to one low mantissa bit
    A2 = np.ascontiguousarray(Am.reshape(m * k, 2)); B2 = np.ascontiguousarray(Bm.reshape(k * n, 2))
    C0 = qgen.matrix(rng, m, n, "D113", n)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_keep(17); qb.set_tensor_scheme(0)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, 1.0, to_dev(A2), k, to_dev(B2), n, 0.0, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert 4 * (n // 2) <= st["flagged"] <= 1024 and st["redo_passes"] == 0, st   # 512 entries <= list capacity: k_oz_fixup, no redo
    _contract_check(qb, oracle, m, n, k, A2, B2, to_host(dC))


def test_bounded_redo_when_many_entries_vanish(qb, oracle):
    """A quarter of the rows of A are zero: every entry of those C rows is an exact 0, far more than 1/64 of the pass, so the
    pass is redone with all diagonals for them; alpha/beta epilogue included (C_in of the flagged entries must be intact)."""
    rng = np.random.default_rng(78)
    m, n, k = 256, 256, 384
    A = qgen.matrix(rng, m, k, "D113", k).reshape(m, k, 2)
    A[::4] = 0
    A = np.ascontiguousarray(A.reshape(m * k, 2)); B = qgen.matrix(rng, k, n, "D113", n); C0 = qgen.matrix(rng, m, n, "D113", n)
    alpha, beta = quad.random_quads(rng, 2)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_keep(17); qb.set_tensor_scheme(0)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, alpha, to_dev(A), k, to_dev(B), n, beta, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["redo_passes"] == 1 and st["flagged"] >= (m // 4) * n, st
    got = to_host(dC).reshape(m, n, 2)
    # the vanished rows: exact sum 0 -> C = fma(alpha, +0, mul(beta, C_in)), bit for bit
    z = np.zeros(((m // 4) * n, 2), dtype=np.uint64)
    want = _epilogue(oracle, alpha, z, beta, np.ascontiguousarray(C0.reshape(m, n, 2)[::4].reshape(-1, 2)))
    assert quad.same_bits(got[::4].reshape(-1, 2), want).all()
    # the other rows went through the bounded path: exact-rounded sums agree to within the contract (sampled)
    s = exact_matmul_rounded(A, k, B, n, m, n, k, "R").reshape(m, n, 2)
    rows = [1, 2, 3, 129, 255]
    for i in rows:
        w = _epilogue(oracle, alpha, np.ascontiguousarray(s[i]), beta, np.ascontiguousarray(C0.reshape(m, n, 2)[i]))
        for j in (0, 100, 255):
            fg, fw = quad.to_fraction(int(got[i, j, 1]), int(got[i, j, 0])), quad.to_fraction(int(w[j, 1]), int(w[j, 0]))
            assert abs(fg - fw) <= abs(fw) / 2 ** 90


# ------------------------------------------------------------------ residue scheme specifics (csrc/qb_crt.cuh)
def test_residue_scheme_row_passes_and_epilogue(qb, oracle):
    """Row passes (pass hook with min_passes = 3) + alpha/beta epilogue + padded leading dimensions: every pass slices its own A
    rows, reuses the residue planes of B, and the result is the exact product rounded once, bit for bit."""
    m, n, k = 384, 264, 200
    rng = np.random.default_rng(11)
    lda, ldb, ldc = k + 3, n + 1, n + 2
    A = qgen.matrix(rng, m, k, "D113", lda); B = qgen.matrix(rng, k, n, "D113", ldb); C0 = qgen.matrix(rng, m, n, "D113", ldc)
    alpha, beta = quad.random_quads(rng, 2)
    rows_idx = rng.choice(m, 6, replace=False)
    seen = []
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_scheme(1)
    try:
        qb.set_gemm_pass_callback(lambda r0, rows: seen.append((r0, rows)), 3)
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_gemm_pass_callback(None)
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["scheme"] == "residues" and st["row_passes"] == 3 and seen == [(0, 128), (128, 128), (256, 128)], (st, seen)
    got = to_host(dC)
    for i in rows_idx:
        s = exact_matmul_rounded(A[i * lda:(i + 1) * lda], lda, B, ldb, 1, n, k, "R")
        want = _epilogue(oracle, alpha, s, beta, C0[i * ldc:i * ldc + n])
        assert quad.same_bits(got[i * ldc:i * ldc + n], want).all(), (i, st)


def test_residue_scheme_zero_operand(qb, oracle):
    """An all-zero A (no span at all): the sums are +0 and C = fma(alpha, +0, mul(beta, C))."""
    m, n, k = 130, 258, 130
    rng = np.random.default_rng(12)
    A = quad.from_double(np.zeros((m, k))).reshape(-1, 2); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)
    zero = quad.from_double(np.zeros(m * n))
    want = _epilogue(oracle, alpha, zero, beta, C0)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_scheme(1)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, alpha, to_dev(A), k, to_dev(B), n, beta, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["scheme"] == "residues", st
    assert quad.same_bits(to_host(dC), want).all()


# ------------------------------------------------------------------ fused gather (qb_set_gemm_peer_outputs): single-GPU check of the store path
@pytest.mark.parametrize("layout,m,n,k", [("R", 200, 250, 300), ("R", 130, 1024, 257), ("C", 140, 130, 260)])
def test_residue_scheme_peer_outputs_mirror_C(qb, oracle, layout, m, n, k):
    """The reconstruction kernel also stores every finished element into the 'peer' copies of C.  On one GPU the peers are simply two
    more buffers of the same device: they must receive exactly the m x n block (same bits as C, padding untouched), through the
    staged 512-byte warp stores for row-major C (ragged n % 4 and n % 128 edges) and the element-wise stores for col-major C."""
    rng = np.random.default_rng(m + n)
    ar, ac = (m, k) if layout == "R" else (k, m)
    br, bc = (k, n) if layout == "R" else (n, k)
    cr, cc = (m, n) if layout == "R" else (n, m)
    lda, ldb, ldc = ac + 1, bc + 2, cc + 5
    A = qgen.matrix(rng, ar, ac, "D113", lda); B = qgen.matrix(rng, br, bc, "D113", ldb); C0 = qgen.matrix(rng, cr, cc, "D113", ldc)
    alpha, beta = quad.random_quads(rng, 2)
    sentinel = qgen.matrix(rng, cr, cc, "D113", ldc)
    dC = to_dev(C0); P1 = to_dev(sentinel.copy()); P2 = to_dev(sentinel.copy())
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_scheme(1)
    try:
        qb.set_gemm_peer_outputs([P1.data_ptr(), P2.data_ptr()])
        qb.gemm(layout, m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc)
        torch.cuda.synchronize()
        wrote = qb.gemm_peer_written()
        # reference order has no fused stores: reports 0 and leaves the peers alone
        qb.set_mode(qb.MODE_REFERENCE)
        P3 = to_dev(sentinel.copy())
        qb.set_gemm_peer_outputs([P3.data_ptr()])
        qb.gemm(layout, 8, 8, 8, alpha, to_dev(A), lda, to_dev(B), ldb, beta, to_dev(C0), ldc)
        torch.cuda.synchronize()
        wrote_ref = qb.gemm_peer_written()
    finally:
        qb.set_gemm_peer_outputs(None)
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert wrote == 2 and wrote_ref == 0
    got = to_host(dC)
    idx = np.array([[(i * ldc + j) if layout == "R" else (j * ldc + i) for j in range(n)] for i in range(m)]).reshape(-1)
    mask = np.ones(C0.shape[0], dtype=bool); mask[idx] = False
    for P in (P1, P2):
        p = to_host(P)
        assert (p[idx] == got[idx]).all(), "peer copy differs from C"
        assert (p[mask] == sentinel[mask]).all(), "peer padding was touched"
    assert (to_host(P3) == sentinel).all()
    assert (got[mask] == C0[mask]).all()


# ------------------------------------------------------------------ more template instances of the residue kernels
def _truncate_mantissa(q, bits):
    """keep the `bits` leading mantissa bits (incl. the implicit one) of every quad: shorter spans, fewer words / moduli"""
    q = q.copy()
    drop = 113 - bits
    if drop >= 64:
        q[..., 0] = 0
        q[..., 1] &= ~np.uint64((1 << (drop - 64)) - 1)
    elif drop > 0:
        q[..., 0] &= ~np.uint64((1 << drop) - 1)
    return q


@pytest.mark.parametrize("name,bits,binades,m,n,k", [("float32-like", 24, 3, 40, 50, 100), ("80-bit", 80, 6, 33, 47, 64), ("96-bit", 96, 8, 20, 300, 96),
                                                    ("wide", 113, 24, 30, 40, 256)])
def test_residue_scheme_other_word_and_group_counts(qb, oracle, name, bits, binades, m, n, k):
    """Spans of ~30 / ~90 / ~110 / ~160 bits: 1, 3, 4 and 6 integer words per element and 2 to 12 reconstruction groups (the
    float32-like case also stands for low-precision data cast to quad).  Exactly rounded, bit for bit."""
    rng = np.random.default_rng(bits + k)
    def mk(r, c):
        return np.ascontiguousarray(_truncate_mantissa(quad.random_quads(rng, (r, c), "D113", emin=-binades, emax=binades), bits).reshape(r * c, 2))
    A = mk(m, k); B = mk(k, n); C0 = mk(m, n)
    alpha, beta = quad.random_quads(rng, 2)
    s = exact_matmul_rounded(A, k, B, n, m, n, k)
    want = _epilogue(oracle, alpha, s, beta, C0)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS); qb.set_tensor_scheme(1)
    try:
        dC = to_dev(C0)
        qb.gemm("R", m, n, k, alpha, to_dev(A), k, to_dev(B), n, beta, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["scheme"] == "residues", st
    assert quad.same_bits(to_host(dC), want).all(), st


def test_residue_scheme_pass_shapes_agree(qb, oracle):
    """qb_set_tensor_pass_shape(1) (short first and last row pass, experimental) changes only WHEN rows are produced: the result is
    bit for bit the one of the equal split, and a few rows around the pass boundaries match exact arithmetic."""
    m, n, k = 4224, 256, 256                      # equal: 1152 x 3 + 768; shaped: 384, 1152 x 3, 384 (api.crt_pass_rows(4224, 1152, s))
    A = dev_random((m * k,), "D113", seed=21); B = dev_random((k * n,), "D113", seed=22); C0 = dev_random((m * n,), "D113", seed=23)
    outs, plans = [], []
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_scheme(1)
    try:
        for shape in (0, 1):
            qb.set_tensor_pass_shape(shape)
            C = C0.clone()
            qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
            torch.cuda.synchronize()
            outs.append(to_host(C)); plans.append(qb.oz_last_stats())
    finally:
        qb.set_tensor_pass_shape(0)
        qb.set_mode(qb.MODE_REFERENCE)
    assert plans[0]["scheme"] == "residues" and plans[0]["row_passes"] == 4 and plans[1]["row_passes"] == 5, plans
    assert (outs[0] == outs[1]).all()
    Ah, Bh = to_host(A), to_host(B)
    one, zero = quad.from_double(np.array([1.0]))[0], quad.from_double(np.array([0.0]))[0]
    for i in (0, 383, 384, 1535, 1536, 3839, 3840, m - 1):
        s = exact_matmul_rounded(Ah[i * k:(i + 1) * k], k, Bh, n, 1, n, k)
        want = _epilogue(oracle, one, s, zero, to_host(C0)[i * n:(i + 1) * n])
        assert quad.same_bits(outs[1][i * n:(i + 1) * n], want).all(), i
