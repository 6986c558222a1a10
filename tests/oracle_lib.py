"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/qoracle.c) and to the reference's
own loops compiled from /root/reference (oracle/_ref/libqref.so).  Only tests/, smoke() and
bench.py's cpu_baseline leg import this module; the product never does."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libqoracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libqref.so")

vp, cl, ci, cc, cd = C.c_void_p, C.c_long, C.c_int, C.c_char, C.c_double


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL)


def _p(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] or a.base is not None
    return C.c_void_p(a.ctypes.data)


def q1(v):
    """scalar -> contiguous (2,) uint64 (callers must hold the result across the foreign call)"""
    from qblas_b200 import quad
    if isinstance(v, (float, int)):
        return np.ascontiguousarray(quad.from_double(np.array([float(v)]))[0])
    return np.ascontiguousarray(np.asarray(v, dtype=np.uint64).reshape(2))


class Oracle:
    def __init__(self, L):
        self.L = L
        L.orc_fma_n.argtypes = [cl, vp, vp, vp, vp]
        L.orc_add_n.argtypes = [cl, vp, vp, vp]
        L.orc_mul_n.argtypes = [cl, vp, vp, vp]
        L.orc_sqrt_n.argtypes = [cl, vp, vp]
        L.orc_dot.argtypes = [cl, vp, cl, vp, cl, ci, vp]
        L.orc_nrm2.argtypes = [cl, vp, cl, ci, vp]
        L.orc_axpy.argtypes = [cl, vp, vp, cl, vp, cl]
        L.orc_gemv.argtypes = [cc, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl]
        L.orc_c_qgemv.argtypes = [cc, cc, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]
        L.orc_gemm.argtypes = [cc, cl, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl, cl]
        L.orc_gemm_trans.argtypes = [cc, cc, cc, cl, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl, cl]
        L.orc_c_qgemm.argtypes = [cc, cc, cc, ci, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]
        L.orc_gemm_sample.argtypes = [cc, cl, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl, cl, cl, vp, vp]
        L.orc_absdot_sample.argtypes = [cc, cl, vp, cl, vp, cl, cl, vp, vp]
        L.orc_exact_dot_check.argtypes = [cc, cl, vp, cl, vp, cl, cl, vp, vp, vp, vp, vp]
        L.orc_to_double.argtypes = [vp]
        L.orc_to_double.restype = cd
        L.orc_from_double.argtypes = [cd, vp]

    # --- scalar ops (vectorised) ---
    def fma(self, a, b, c):
        out = np.empty_like(a)
        self.L.orc_fma_n(a.size // 2, _p(a), _p(b), _p(c), _p(out))
        return out

    def add(self, a, b):
        out = np.empty_like(a)
        self.L.orc_add_n(a.size // 2, _p(a), _p(b), _p(out))
        return out

    def mul(self, a, b):
        out = np.empty_like(a)
        self.L.orc_mul_n(a.size // 2, _p(a), _p(b), _p(out))
        return out

    def sqrt(self, a):
        out = np.empty_like(a)
        self.L.orc_sqrt_n(a.size // 2, _p(a), _p(out))
        return out

    def to_double(self, q):
        qq = np.ascontiguousarray(q)
        return self.L.orc_to_double(_p(qq))

    def from_double(self, d):
        out = np.zeros(2, dtype=np.uint64)
        self.L.orc_from_double(float(d), _p(out))
        return out

    # --- routines (operate in place on y / C like the reference) ---
    def dot(self, n, x, incx, y, incy, T):
        out = np.zeros(2, dtype=np.uint64)
        self.L.orc_dot(n, _p(x), incx, _p(y), incy, T, _p(out))
        return out

    def nrm2(self, n, x, incx, T):
        out = np.zeros(2, dtype=np.uint64)
        self.L.orc_nrm2(n, _p(x), incx, T, _p(out))
        return out

    def axpy(self, n, alpha, x, incx, y, incy):
        al = q1(alpha)  # keep alive across the call
        self.L.orc_axpy(n, _p(al), _p(x), incx, _p(y), incy)

    def gemv(self, layout, m, n, alpha, A, lda, x, incx, beta, y, incy):
        al, be = q1(alpha), q1(beta)
        self.L.orc_gemv(layout.encode(), m, n, _p(al), _p(A), lda, _p(x), incx, _p(be), _p(y), incy)

    def c_qgemv(self, layout, trans, m, n, alpha, A, lda, x, incx, beta, y, incy):
        self.L.orc_c_qgemv(layout.encode(), trans.encode(), m, n, alpha, _p(A), lda, _p(x), incx, beta, _p(y), incy)

    def gemm(self, layout, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, kc=126):
        al, be = q1(alpha), q1(beta)
        self.L.orc_gemm(layout.encode(), m, n, k, _p(al), _p(A), lda, _p(B), ldb, _p(be), _p(Cm), ldc, kc)

    def gemm_trans(self, layout, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, kc=126):
        al, be = q1(alpha), q1(beta)
        self.L.orc_gemm_trans(layout.encode(), ta.encode(), tb.encode(), m, n, k, _p(al), _p(A), lda, _p(B), ldb,
                              _p(be), _p(Cm), ldc, kc)

    def c_qgemm(self, layout, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        self.L.orc_c_qgemm(layout.encode(), ta.encode(), tb.encode(), m, n, k, alpha, _p(A), lda, _p(B), ldb, beta, _p(Cm), ldc)

    def gemm_sample(self, layout, m, n, k, alpha, A, lda, B, ldb, beta, Cin, ldc, idx, kc=126):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.zeros((idx.shape[0], 2), dtype=np.uint64)
        al, be = q1(alpha), q1(beta)
        self.L.orc_gemm_sample(layout.encode(), m, n, k, _p(al), _p(A), lda, _p(B), ldb, _p(be),
                               _p(Cin) if Cin is not None else None, ldc, kc, idx.shape[0], C.c_void_p(idx.ctypes.data), _p(out))
        return out

    def absdot_sample(self, layout, k, A, lda, B, ldb, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.zeros((idx.shape[0], 2), dtype=np.uint64)
        self.L.orc_absdot_sample(layout.encode(), k, _p(A), lda, _p(B), ldb, idx.shape[0], C.c_void_p(idx.ctypes.data), _p(out))
        return out


    def exact_dot_check(self, layout, k, A, lda, B, ldb, idx, got=None):
        """Exact inner products of the sampled (i, j) through a long accumulator (oracle/qoracle.c): returns
        (exact sums rounded once (ns, 2), err / (k u sum|a||b|) per entry as float64 (<= 1 <=> fast-mode contract), IEEE class per entry:
        0 finite data, 1 NaN, 2 +Inf, 3 -Inf).  `got`: the (ns, 2) results to check (None: only the exact sums)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        ns = idx.shape[0]
        exact = np.zeros((ns, 2), dtype=np.uint64)
        ratio = np.zeros(ns, dtype=np.float64)
        klass = np.zeros(ns, dtype=np.int32)
        if got is not None:
            got = np.ascontiguousarray(got, dtype=np.uint64)
            assert got.shape == (ns, 2)
        self.L.orc_exact_dot_check(layout.encode(), k, _p(A), lda, _p(B), ldb, ns, C.c_void_p(idx.ctypes.data),
                                   _p(got) if got is not None else None, _p(exact), C.c_void_p(ratio.ctypes.data), C.c_void_p(klass.ctypes.data))
        return exact, ratio, klass


class Ref:
    """The reference's own loops (unmodified headers + libquadmath shim)."""

    def __init__(self, L):
        self.L = L
        L.ref_fma_n.argtypes = [cl, vp, vp, vp, vp]
        L.ref_gemm.argtypes = [cc, cl, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl]
        L.ref_gemv.argtypes = [cc, cl, cl, vp, vp, cl, vp, cl, vp, vp, cl]
        L.ref_dot.argtypes = [cl, vp, cl, vp, cl, vp]
        L.ref_nrm2.argtypes = [cl, vp, cl, vp]
        L.ref_axpy.argtypes = [cl, vp, vp, cl, vp, cl]
        L.ref_c_qdot.argtypes = [ci, vp, ci, vp, ci]
        L.ref_c_qdot.restype = cd
        L.ref_c_qnrm2.argtypes = [ci, vp, ci]
        L.ref_c_qnrm2.restype = cd
        L.ref_c_qaxpy.argtypes = [ci, cd, vp, ci, vp, ci]
        L.ref_c_qgemv.argtypes = [cc, cc, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]
        L.ref_c_qgemm.argtypes = [cc, cc, cc, ci, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]
        L.ref_set_num_threads.argtypes = [ci]
        L.ref_get_num_threads.restype = ci
        L.ref_get_version.restype = C.c_char_p
        L.ref_arith.restype = C.c_char_p

    def set_num_threads(self, t):
        self.L.ref_set_num_threads(t)

    def fma(self, a, b, c):
        out = np.empty_like(a)
        self.L.ref_fma_n(a.size // 2, _p(a), _p(b), _p(c), _p(out))
        return out

    def gemm(self, layout, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        al, be = q1(alpha), q1(beta)
        self.L.ref_gemm(layout.encode(), m, n, k, _p(al), _p(A), lda, _p(B), ldb, _p(be), _p(Cm), ldc)

    def gemv(self, layout, m, n, alpha, A, lda, x, incx, beta, y, incy):
        al, be = q1(alpha), q1(beta)
        self.L.ref_gemv(layout.encode(), m, n, _p(al), _p(A), lda, _p(x), incx, _p(be), _p(y), incy)

    def dot(self, n, x, incx, y, incy):
        out = np.zeros(2, dtype=np.uint64)
        self.L.ref_dot(n, _p(x), incx, _p(y), incy, _p(out))
        return out

    def nrm2(self, n, x, incx):
        out = np.zeros(2, dtype=np.uint64)
        self.L.ref_nrm2(n, _p(x), incx, _p(out))
        return out

    def axpy(self, n, alpha, x, incx, y, incy):
        al = q1(alpha)
        self.L.ref_axpy(n, _p(al), _p(x), incx, _p(y), incy)

    def c_qdot(self, n, x, incx, y, incy):
        return self.L.ref_c_qdot(n, _p(x), incx, _p(y), incy)

    def c_qnrm2(self, n, x, incx):
        return self.L.ref_c_qnrm2(n, _p(x), incx)

    def c_qgemv(self, layout, trans, m, n, alpha, A, lda, x, incx, beta, y, incy):
        self.L.ref_c_qgemv(layout.encode(), trans.encode(), m, n, alpha, _p(A), lda, _p(x), incx, beta, _p(y), incy)

    def c_qgemm(self, layout, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        self.L.ref_c_qgemm(layout.encode(), ta.encode(), tb.encode(), m, n, k, alpha, _p(A), lda, _p(B), ldb, beta, _p(Cm), ldc)


_oracle = None
_ref = None


def load_oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "qoracle.c")):
            build()
        _oracle = Oracle(C.CDLL(ORACLE_SO))
    return _oracle


def load_ref():
    """None when neither /root/reference nor a prebuilt oracle/_ref/libqref.so exists."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/include/quadblas"):
            build()
        if not os.path.exists(REF_SO):
            return None
        _ref = Ref(C.CDLL(REF_SO))
    return _ref
