"""GPU tests of the fast-mode tensor-core qgemm (csrc/qb_ozaki.cu + csrc/qb_crt.cuh): the int8 tcgen05 kernel alone against
integer matmul, and the whole path (scan -> residues -> mma -> reconstruction [-> fix-up]) against EXACT inner products rounded once
(tests/exact_ref.py in Python integers for small cases, the long accumulator of oracle/qoracle.c for large ones) followed by the
reference epilogue C = fma(alpha, s, mul(beta, C)) (/root/reference/include/quadblas/algorithms/level3.hpp:102-109).  Inputs whose
exponent spread the moduli cannot cover (SURVEY.md §8d cfg3 'Dexp') are checked against the fast-mode contract
|c^ - c| <= gamma_k (|A||B|)_ij entry by entry."""
import numpy as np
import pytest
import torch

import qgen
from exact_ref import exact_matmul_rounded
from gpu_util import dev_random, to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True)
def _defaults(qb):
    yield
    qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO); qb.set_tensor_unit(0, 0); qb.set_tensor_window(144)
    qb.set_tensor_workspace_limit(0); qb.set_tensor_ramp(0, 0)
    qb.set_gemm_pass_callback(None); qb.set_gemm_peer_outputs(None); qb.set_gemm_b_panels(None); qb.set_gemm_b_planes(0)


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



def _mk(rng, r, c, ld, kind):
    if kind.startswith("Dexp"):  # full mantissas times 2^U{-s..s}
        s_ = int(kind[4:])
        return np.ascontiguousarray(quad.random_quads(rng, (r, ld), "D113", emin=-s_, emax=s_).reshape(r * ld, 2))
    if kind == "Dint":   # small integers (a few moduli, one reconstruction group), with zeros
        return quad.from_double(rng.integers(-9, 10, size=(r, ld)).astype(np.float64)).reshape(r * ld, 2)
    return qgen.matrix(rng, r, c, kind, ld)


def _fast_gemm(qb, layout, m, n, k, alpha, A, lda, B, ldb, beta, C0, ldc):
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
    dC = to_dev(C0)
    qb.gemm(layout, m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc)
    torch.cuda.synchronize()
    st = qb.oz_last_stats()
    qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert st["pairs"] > 0, "tensor path declined"
    return to_host(dC), st


@pytest.mark.parametrize("m,n,k,kind,layout", [(40, 33, 300, "D113", "R"), (33, 20, 257, "D53", "C"), (130, 260, 64, "D113", "R"), (5, 7, 1000, "D113", "C"),
                                               (24, 30, 140, "Dexp8", "R"), (17, 50, 129, "Dexp12", "C"), (9, 300, 70, "Dint", "C"), (140, 12, 33, "D113", "C")])
def test_fast_gemm_tensor_path_is_exactly_rounded(qb, oracle, m, n, k, kind, layout):
    """Spans the moduli cover (everything up to +-12 binades inside a row): exact inner products, one rounding, bit for bit."""
    rng = np.random.default_rng(m + n + k)
    ar, ac = (m, k) if layout == "R" else (k, m)
    br, bc = (k, n) if layout == "R" else (n, k)
    cr, cc = (m, n) if layout == "R" else (n, m)
    lda, ldb, ldc = ac + 1, bc + 2, cc + 3
    A = _mk(rng, ar, ac, lda, kind); B = _mk(rng, br, bc, ldb, kind); C0 = _mk(rng, cr, cc, ldc, kind)
    alpha, beta = quad.random_quads(rng, 2)
    s = exact_matmul_rounded(A, lda, B, ldb, m, n, k, layout)          # (m*n, 2) in (i, j) order
    idx = np.array([[(i * ldc + j) if layout == "R" else (j * ldc + i) for j in range(n)] for i in range(m)]).reshape(-1)
    want = _epilogue(oracle, alpha, s, beta, C0[idx])
    got, st = _fast_gemm(qb, layout, m, n, k, alpha, A, lda, B, ldb, beta, C0, ldc)
    assert st["exact"] and st["flagged"] == 0, st
    assert quad.same_bits(got[idx], want).all(), f"{(~quad.same_bits(got[idx], want)).sum()} mismatches, plan {st}"
    mask = np.ones(C0.shape[0], dtype=bool); mask[idx] = False       # untouched padding
    assert (got[mask] == C0[mask]).all()


def test_fast_gemm_cancellation_pattern_is_exact(qb, oracle):
    """README:141-158 / test_quadblas.cpp:715-739: rows (1e20, 1, -1e20, 0...) times ones -> exactly 1."""
    m, n, k = 128, 128, 256
    A = np.zeros((m, k)); A[:, 0] = 1e20; A[:, 1] = 1.0; A[:, 2] = -1e20
    Aq = quad.from_double(A).reshape(-1, 2); Bq = quad.from_double(np.ones((k, n))).reshape(-1, 2)
    C0 = quad.from_double(np.zeros((m, n))).reshape(-1, 2)
    qb.set_mode(qb.MODE_FAST)
    dC = to_dev(C0)
    qb.gemm("R", m, n, k, 1.0, to_dev(Aq), k, to_dev(Bq), n, 0.0, dC, n)
    torch.cuda.synchronize()
    st = qb.oz_last_stats()
    assert st["pairs"] > 0 and st["exact"], st
    assert quad.same_bits(to_host(dC), quad.from_double(np.ones(m * n))).all()


def _check_contract(oracle, layout, m, n, k, A, lda, B, ldb, got_ij, idx=None):
    """every (sampled) entry against the exact inner product: returns (exact sums rounded once, err / (k u |A||B|), IEEE class)"""
    if idx is None:
        idx = np.stack(np.meshgrid(np.arange(m), np.arange(n), indexing="ij"), axis=-1).reshape(-1, 2)
    return oracle.exact_dot_check(layout, k, A, lda, B, ldb, idx, got_ij)


@pytest.mark.parametrize("m,n,k,kind,window", [(256, 256, 1024, "Dexp40", 144), (130, 260, 300, "Dexp40", 144), (128, 256, 512, "Dexp40", 128),
                                               (140, 132, 260, "Dexp100", 144)])
def test_wide_exponent_ranges_stay_on_the_tensor_path_and_meet_the_contract(qb, oracle, m, n, k, kind, window):
    """SURVEY.md §8d cfg3 'Dexp' = D113 x 2^U{-40..40} (and wider): row spans of ~190+ bits, more than the moduli cover.  The planner
    caps the windows instead of declining, the reconstruction kernel tests every element against the dropped mass and the fix-up
    kernel recomputes the rejected ones.  EVERY entry of C is inside |c^ - c| <= k u (|A||B|)_ij against exact arithmetic."""
    rng = np.random.default_rng(m * 3 + k)
    A = _mk(rng, m, k, k, kind); B = _mk(rng, k, n, n, kind); C0 = _mk(rng, m, n, n, "D113")
    qb.set_tensor_window(window)
    got, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    assert st["truncated"] & 3 and not st["exact"] and st["WA"] + st["WB"] <= 2 * window, st
    exact, ratio, klass = _check_contract(oracle, "R", m, n, k, A, k, B, n, got)
    assert (klass == 0).all() and (ratio <= 1.0).all(), (st, float(ratio.max()), int((ratio > 1).sum()))
    if kind == "Dexp40" and window == 144 and k >= 1024:
        assert st["flagged"] <= m * n // 1000, st          # a handful of cancelling entries at most
    # deterministic
    got2, _ = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    assert (got == got2).all()


def test_cancelling_entries_under_a_capped_window_go_to_the_fixup(qb, oracle):
    """Rows whose large terms cancel exactly (+2^60 b, -2^60 b): what is left comes from small elements a capped window truncates; those
    entries fail the acceptance test and are left to k_crt_fixup (window accumulator): inside the contract; alpha / beta epilogue
    included (C_in of a rejected entry must still be intact when the fix-up reads it)."""
    rng = np.random.default_rng(77)
    m, n, k = 128, 256, 512
    A = _mk(rng, m, k, k, "Dexp40").reshape(m, k, 2); B = _mk(rng, k, n, n, "Dexp40").reshape(k, n, 2)
    big = quad.from_double(np.array([2.0 ** 60]))[0]
    A[:8, 0] = big; A[:8, 1] = big; A[:8, 1, 1] ^= np.uint64(1 << 63)
    A[:8, 2:, 1] = (A[:8, 2:, 1] & np.uint64(0x8000ffffffffffff)) | (np.uint64(16383 - 30) << np.uint64(48))    # the rest of those rows is small
    B[1, :] = B[0, :]
    A = np.ascontiguousarray(A.reshape(m * k, 2)); B = np.ascontiguousarray(B.reshape(k * n, 2)); C0 = _mk(rng, m, n, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)
    got, st = _fast_gemm(qb, "R", m, n, k, alpha, A, k, B, n, beta, C0, n)
    assert st["truncated"] & 3 and st["flagged"] >= 8 * n, st
    # the fix-up's window accumulator keeps 192 bits below the largest product of the row (the cancelled +-2^60 b pair here), which leaves
    # ~80 correct bits for sums 2^-90 of it: inside the contract, not bit-exact.  alpha / beta epilogue: applied to the same sums.
    got0, st0 = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    exact, ratio, klass = _check_contract(oracle, "R", m, n, k, A, k, B, n, got0)
    assert (klass == 0).all() and (ratio <= 1.0).all(), float(ratio.max())
    assert ratio[:8 * n].max() < 1e-3                       # the fixed-up rows are far inside the bound
    want = _epilogue(oracle, alpha, got0, beta, C0)          # same sums through the reference epilogue, bit for bit
    assert quad.same_bits(got, want).all()


def test_inf_nan_and_huge_spans_run_on_the_tensor_path(qb, oracle):
    """Inf / NaN operands and a row that spans 2000 binades: nothing is declined any more.  The rows / columns that hold Inf or NaN get
    the IEEE result of the sum of products (NaN for NaN, Inf - Inf, Inf * 0; else +-Inf) from the fix-up kernel, every finite entry
    meets the contract, and the rest of the matrix is not slowed down to the integer kernel."""
    rng = np.random.default_rng(3)
    m, n, k = 130, 140, 260
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    inf, nan = quad.from_double(np.array([np.inf]))[0], quad.from_double(np.array([np.nan]))[0]
    A[5] = inf                                   # row 0: +Inf (times finite non-zero)
    A[2 * k + 7] = inf; A[2 * k + 8] = inf; A[2 * k + 8, 1] ^= np.uint64(1 << 63)     # row 2: +Inf and -Inf (NaN where the signs of b make them oppose)
    A[4 * k + 1] = nan                           # row 4: NaN
    B[9 * n + 3] = inf                           # column 3: Inf
    B[11 * n + 6] = quad.from_double(np.array([0.0]))[0]; A[7 * k + 11] = inf          # row 7, column 6: Inf * 0
    A[9 * k + 0] = quad.from_double(np.array([1e-300]))[0]; A[9 * k + 1] = quad.from_double(np.array([1e300]))[0]   # row 9: 2000 binades
    got, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    assert st["truncated"] & 4 and st["flagged"] >= 4 * n + m - 4, st
    exact, ratio, klass = _check_contract(oracle, "R", m, n, k, A, k, B, n, got)
    assert (ratio <= 1.0).all(), (st, np.argwhere(ratio > 1)[:5])
    kl = klass.reshape(m, n)
    assert (kl[4] == 1).all() and (kl[2] != 0).all() and (kl[2] == 1).any() and kl[7, 6] == 1 and (kl[:, 3] != 0).all() and (kl[0] != 0).all()
    assert (kl[[1, 3, 5, 6, 8, 9] + list(range(10, m))][:, [0, 1, 2] + list(range(4, n))] == 0).all()
    fin = (klass == 0).reshape(m, n).copy(); fin[9] = False     # row 9 lost bits to its window: contract only; all others are exact sums
    assert quad.same_bits(got.reshape(m, n, 2)[fin], exact.reshape(m, n, 2)[fin]).all()


@pytest.mark.parametrize("unit", [(128, 256), (256, 512), (384, 256)])
def test_units_passes_panels_and_epilogue(qb, oracle, unit):
    """The pipeline units (row passes x column panels): small units force many passes and panels, ragged last ones included; the pass
    hook reports every row once; alpha/beta epilogue and padded leading dimensions; sampled rows against exact arithmetic, and the
    whole result bit for bit the one of a single unit."""
    m, n, k = 520, 700, 200
    rng = np.random.default_rng(11)
    lda, ldb, ldc = k + 3, n + 1, n + 2
    A = qgen.matrix(rng, m, k, "D113", lda); B = qgen.matrix(rng, k, n, "D113", ldb); C0 = qgen.matrix(rng, m, n, "D113", ldc)
    alpha, beta = quad.random_quads(rng, 2)
    got1, st1 = _fast_gemm(qb, "R", m, n, k, alpha, A, lda, B, ldb, beta, C0, ldc)
    assert st1["units"] == 1
    seen = []
    qb.set_tensor_unit(*unit)
    qb.set_gemm_pass_callback(lambda r0, rows: seen.append((r0, rows)), 1)
    got, st = _fast_gemm(qb, "R", m, n, k, alpha, A, lda, B, ldb, beta, C0, ldc)
    qb.set_gemm_pass_callback(None)
    np_, nj = -(-m // unit[0]), -(-n // unit[1])
    assert st["row_passes"] == np_ and st["panels"] == nj and st["units"] == np_ * nj, st
    assert sorted(seen) == [(p * unit[0], min(unit[0], m - p * unit[0])) for p in range(np_)], seen
    assert (got == got1).all()
    for i in rng.choice(m, 5, replace=False):
        s = exact_matmul_rounded(A[i * lda:(i + 1) * lda], lda, B, ldb, 1, n, k, "R")
        want = _epilogue(oracle, alpha, s, beta, C0[i * ldc:i * ldc + n])
        assert quad.same_bits(got[i * ldc:i * ldc + n], want).all(), (i, st)


@pytest.mark.parametrize("host", [False, True])
def test_workspace_limit_sweeps_in_blocks(qb, oracle, host):
    """A capped workspace (qb_set_tensor_workspace_limit): only some passes (device path: panels outer) or some panels (all-host path:
    passes outer) stay resident, the product is swept in blocks, the slots of the other operand are reused under their events.  Same bits
    as the unconstrained call; a cap too small for a single unit makes the planner decline (integer-limb kernel, contract only)."""
    m, n, k = (1664, 1536, 4608) if host else (1100, 1300, 512)         # host: >= 64 MB in total for the pipelined path
    rng = np.random.default_rng(31 + host)
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)

    def run():
        qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
        l0 = qb.launch_count()
        if host:
            C = C0.copy()
            qb.gemm("R", m, n, k, alpha, A, k, B, n, beta, C, n)
        else:
            dC = to_dev(C0)
            qb.gemm("R", m, n, k, alpha, to_dev(A), k, to_dev(B), n, beta, dC, n)
            torch.cuda.synchronize()
            C = to_host(dC)
        st = qb.oz_last_stats()
        st["launches"] = qb.launch_count() - l0
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
        return C, st

    want, st0 = run()
    assert st0["pairs"] > 0
    qb.set_tensor_unit(256, 256)
    full, st1 = run()
    assert st1["units"] == st1["row_passes"] * st1["panels"] and (full == want).all()
    N, Kp = st1["pairs"], st1["Kp"]
    # room for 2 + 2 slots of 256 rows and the two residue buffers, far less than all passes / panels resident
    qb.set_tensor_workspace_limit(5 * N * 256 * Kp + 2 * N * (256 * 256 + 4352) + (64 << 10))
    got, st2 = run()
    assert st2["pairs"] > 0 and st2["units"] == st1["units"] and st2["launches"] > st1["launches"], (st1, st2)   # blocks: the residues of the non-resident operand are recomputed
    assert (got == want).all()
    qb.set_tensor_workspace_limit(1 << 20)
    low, st3 = run()
    assert st3["pairs"] == 0                                                 # declined -> integer-limb kernel
    idx = np.stack([rng.integers(0, m, 64), rng.integers(0, n, 64)], axis=1)
    qb.set_tensor_workspace_limit(0)
    pure, _ = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n) if not host else (None, None)
    if not host:
        _, ratio, _ = oracle.exact_dot_check("R", k, A, k, B, n, idx, pure.reshape(m, n, 2)[idx[:, 0], idx[:, 1]])
        assert (ratio <= 1.0).all()


def test_pass_callback_min_passes(qb, oracle):
    m, n, k = 384, 264, 200
    rng = np.random.default_rng(12)
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    seen = []
    qb.set_gemm_pass_callback(lambda r0, rows: seen.append((r0, rows)), 3)
    got, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    qb.set_gemm_pass_callback(None)
    assert st["row_passes"] == 3 and seen == [(0, 128), (128, 128), (256, 128)], (st, seen)
    # min_passes larger than ceil(m / 128): passes of one tile row each (include/qblas_b200.h)
    seen.clear()
    qb.set_gemm_pass_callback(lambda r0, rows: seen.append((r0, rows)), 16)
    got2, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    qb.set_gemm_pass_callback(None)
    assert len(seen) == 3 and (got == got2).all()


def test_zero_operand(qb, oracle):
    """An all-zero A (no span at all): the sums are +0 and C = fma(alpha, +0, mul(beta, C))."""
    m, n, k = 130, 258, 130
    rng = np.random.default_rng(12)
    A = quad.from_double(np.zeros((m, k))).reshape(-1, 2); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)
    want = _epilogue(oracle, alpha, quad.from_double(np.zeros(m * n)), beta, C0)
    got, st = _fast_gemm(qb, "R", m, n, k, alpha, A, k, B, n, beta, C0, n)
    assert st["exact"] and quad.same_bits(got, want).all(), st


def test_fast_gemm_large_matches_sampled_exact(qb, oracle):
    """1024 x 768 x 2304 with the default units, device resident: sampled entries against the long accumulator, bit for bit."""
    m, n, k = 1024, 768, 2304
    A = dev_random((m * k,), "D113", seed=1); B = dev_random((k * n,), "D113", seed=2); C = dev_random((m * n,), "D113", seed=3)
    qb.set_mode(qb.MODE_FAST)
    qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
    torch.cuda.synchronize()
    st = qb.oz_last_stats()
    assert st["pairs"] > 0 and st["exact"], st
    rng = np.random.default_rng(0)
    idx = np.stack([rng.integers(0, m, 256), rng.integers(0, n, 256)], axis=1)
    got = to_host(C).reshape(m, n, 2)[idx[:, 0], idx[:, 1]]
    exact, ratio, klass = oracle.exact_dot_check("R", k, to_host(A), k, to_host(B), n, idx, got)
    assert quad.same_bits(got, exact).all() and (klass == 0).all()


def test_k_beyond_one_int32_accumulation(qb, oracle):
    """k > 65536: the int32 accumulators of the tensor kernel hold 2^30 at most, so K is cut into chunks whose residues are added
    mod p in the kernel's epilogue.  Sampled entries against the long accumulator, bit for bit."""
    m, n, k = 128, 256, 65536 + 4096 + 100
    A = dev_random((m * k,), "D113", seed=31); B = dev_random((k * n,), "D113", seed=32); C = dev_random((m * n,), "D113", seed=33)
    qb.set_mode(qb.MODE_FAST)
    qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
    torch.cuda.synchronize()
    st = qb.oz_last_stats()
    assert st["pairs"] > 0 and st["nchunks"] == 2 and st["exact"], st
    rng = np.random.default_rng(1)
    idx = np.stack([rng.integers(0, m, 48), rng.integers(0, n, 48)], axis=1)
    got = to_host(C).reshape(m, n, 2)[idx[:, 0], idx[:, 1]]
    exact, _, klass = oracle.exact_dot_check("R", k, to_host(A), k, to_host(B), n, idx, got)
    assert quad.same_bits(got, exact).all() and (klass == 0).all()


def test_streamed_b_panels(qb, oracle):
    """qb_set_gemm_b_panels: B is handed over in column panels that live in their own packed buffers (what a rank that receives B panel
    by panel holds), with the column statistics computed up front; same bits as the resident-B call."""
    m, n, k, pw = 300, 1024, 260, 256
    rng = np.random.default_rng(21)
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)
    want, st0 = _fast_gemm(qb, "R", m, n, k, alpha, A, k, B, n, beta, C0, n)
    dB = to_dev(B)
    stats = torch.zeros(3 * n, dtype=torch.int32, device="cuda")
    qb.gemm_colstats("R", k, n, dB, n, stats)
    panels = [dB.reshape(k, n, 2)[:, c0:c0 + pw].contiguous() for c0 in range(0, n, pw)]     # packed k x pw panels, ld = pw
    calls = []

    def provide(col0, cols, stream_ptr):
        calls.append((col0, cols))
        ev = torch.cuda.Event(); ev.record()                       # "arrival": ordered after the packing copies on torch's stream
        torch.cuda.ExternalStream(stream_ptr).wait_event(ev)
        return panels[col0 // pw].data_ptr(), pw

    qb.set_gemm_b_panels(provide, pw, stats)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
    dC = to_dev(C0)
    bogus = torch.zeros(16, dtype=torch.int64, device="cuda")     # the B argument is not dereferenced
    qb.gemm("R", m, n, k, alpha, to_dev(A), k, bogus.reshape(-1, 2), n, beta, dC, n)
    torch.cuda.synchronize()
    st = qb.oz_last_stats()
    qb.set_gemm_b_panels(None)
    assert calls == [(c0, pw) for c0 in range(0, n, pw)] and st["panels"] == n // pw, (calls, st)
    assert (to_host(dC) == want).all()


def test_b_handed_over_as_residue_planes(qb, oracle):
    """qb_crt_plan + qb_crt_residues_dev + qb_set_gemm_b_planes: the residue planes of B are computed OUTSIDE the call (column slice by
    column slice into one [N][n][Kp] buffer, as the ranks of a row-sharded product do before exchanging them) and handed over panel by
    panel; the call computes no residues of B.  Same bits as the ordinary call; more planes than the call needs are fine; a call that
    would need the fix-up (B's original elements) is refused."""
    m, n, k, pw = 300, 1280, 260, 256
    rng = np.random.default_rng(23)
    A = qgen.matrix(rng, m, k, "D113"); B = qgen.matrix(rng, k, n, "D113"); C0 = qgen.matrix(rng, m, n, "D113")
    alpha, beta = quad.random_quads(rng, 2)
    want, st0 = _fast_gemm(qb, "R", m, n, k, alpha, A, k, B, n, beta, C0, n)
    dA, dB = to_dev(A), to_dev(B)
    statsB = torch.zeros(3 * n, dtype=torch.int32, device="cuda"); statsA = torch.zeros(3 * m, dtype=torch.int32, device="cuda")
    qb.gemm_colstats("R", k, n, dB, n, statsB)
    qb.gemm_colstats("C", k, m, dA, k, statsA)            # the rows of a row-major A are the columns of the col-major view
    span = lambda st, cnt: int(torch.where(st[:cnt] != 0, st[:cnt] + 113 - st[cnt:2 * cnt], torch.zeros_like(st[:cnt])).max().item())
    WAs, WBs = span(statsA, m), span(statsB, n)
    assert (WAs, WBs) == (st0["WA_span"], st0["WB_span"])
    N, WA, WB, trunc = qb.crt_plan(WAs, WBs, k)
    assert trunc == 0 and (N, WA, WB) == (st0["pairs"], st0["WA"], st0["WB"])
    Kp = (k + 127) // 128 * 128
    NB = N + 2                                            # two planes more than this call needs
    planes = torch.zeros((NB, n, Kp), dtype=torch.int8, device="cuda")
    for c0 in range(0, n, 320):                           # slices of 320 columns, written straight into the shared layout
        qb.crt_residues("R", k, 320, None, n, statsB[c0:], WB, NB, planes.data_ptr() + c0 * Kp, n * Kp, data_ptr=dB.data_ptr() + c0 * 16)
    calls = []

    def provide(col0, cols, stream_ptr):
        calls.append(col0)
        ev = torch.cuda.Event(); ev.record()
        torch.cuda.ExternalStream(stream_ptr).wait_event(ev)
        return planes.data_ptr() + col0 * Kp, n * Kp

    qb.set_gemm_b_panels(provide, pw, statsB); qb.set_gemm_b_planes(NB, WB)
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
    l0 = qb.launch_count()
    try:
        dC = to_dev(C0)
        bogus = torch.zeros(16, dtype=torch.int64, device="cuda")
        qb.gemm("R", m, n, k, alpha, dA, k, bogus.reshape(-1, 2), n, beta, dC, n)
        torch.cuda.synchronize()
        st = qb.oz_last_stats()
        launches = qb.launch_count() - l0
        # Dexp rows need the fix-up: refused in this mode
        Ax = to_dev(_mk(rng, m, k, k, "Dexp40"))
        with pytest.raises(qb.QblasError):
            qb.gemm("R", m, n, k, alpha, Ax, k, bogus.reshape(-1, 2), n, beta, to_dev(C0), n)
    finally:
        qb.set_gemm_b_planes(0); qb.set_gemm_b_panels(None)
        qb.set_mode(qb.MODE_REFERENCE); qb.set_tensor_path(qb.TENSOR_AUTO)
    assert calls == list(range(0, n, pw)) and st["pairs"] == N
    assert (to_host(dC) == want).all()


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
    qb.set_mode(qb.MODE_FAST); qb.set_tensor_path(qb.TENSOR_ALWAYS)
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
                                                    ("wide", 113, 14, 30, 40, 256)])
def test_other_word_and_group_counts(qb, oracle, name, bits, binades, m, n, k):
    """Spans of ~30 / ~90 / ~110 / ~140 bits: 1, 3, 4 and 5 integer words per element and 2 to 11 reconstruction groups (the
    float32-like case also stands for low-precision data cast to quad).  Exactly rounded, bit for bit."""
    rng = np.random.default_rng(bits + k)
    def mk(r, c):
        return np.ascontiguousarray(_truncate_mantissa(quad.random_quads(rng, (r, c), "D113", emin=-binades, emax=binades), bits).reshape(r * c, 2))
    A = mk(m, k); B = mk(k, n); C0 = mk(m, n)
    alpha, beta = quad.random_quads(rng, 2)
    s = exact_matmul_rounded(A, k, B, n, m, n, k)
    want = _epilogue(oracle, alpha, s, beta, C0)
    got, st = _fast_gemm(qb, "R", m, n, k, alpha, A, k, B, n, beta, C0, n)
    assert st["exact"], st
    assert quad.same_bits(got, want).all(), st


def test_widest_windows_six_words_thirteen_groups(qb, oracle):
    """A wide A (+-36 binades: 185-bit rows) against double-valued B (54 bits): the narrow operand leaves its share of the moduli to
    the wide one, which takes all 6 words; with qb_set_tensor_window(170) both operands at +-24 binades use 161-bit windows and 13
    reconstruction groups.  Exact in both cases."""
    rng = np.random.default_rng(5)
    m, n, k = 30, 40, 128
    A = _mk(rng, m, k, k, "Dexp36"); B = qgen.matrix(rng, k, n, "D53"); C0 = _mk(rng, m, n, n, "D113")
    got, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    assert st["exact"] and st["WA"] > 160, st
    assert quad.same_bits(got, exact_matmul_rounded(A, k, B, n, m, n, k)).all()
    qb.set_tensor_window(170)
    A = _mk(rng, m, k, k, "Dexp24"); B = _mk(rng, k, n, n, "Dexp24")
    got, st = _fast_gemm(qb, "R", m, n, k, 1.0, A, k, B, n, 0.0, C0, n)
    assert st["exact"] and st["pairs"] >= 45, st
    assert quad.same_bits(got, exact_matmul_rounded(A, k, B, n, m, n, k)).all()
