"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libqref.so = the unmodified
reference headers compiled from /root/reference against the libquadmath shim).  Run in the dev
container (where /root/reference exists):   python tests/golden/make_golden.py
The fixtures are small, committed, and let both the oracle (CPU tests) and the CUDA path (GPU
tests) be checked against reference outputs on a box where /root/reference does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
import qgen  # noqa: E402
from qblas_b200 import quad  # noqa: E402


def main():
    ref = oracle_lib.load_ref()
    assert ref is not None, "needs /root/reference (or a prebuilt oracle/_ref/libqref.so)"
    rng = np.random.default_rng(20261017)
    out = {}
    # scalar fma vectors, all regimes (Sleef_fmaq1_u05 as the reference calls it)
    for reg in qgen.REGIMES:
        a, b, c = qgen.triples(rng, 256, reg)
        out[f"fma_{reg}_a"], out[f"fma_{reg}_b"], out[f"fma_{reg}_c"] = a, b, c
        out[f"fma_{reg}_out"] = ref.fma(a, b, c)
    # gemm row-major (blocked path incl. k-panels of 126) and col-major (simple path, dims <= 64)
    cases = []
    for gi, (lay, m, n, k) in enumerate([("R", 70, 9, 300), ("R", 65, 66, 127), ("R", 33, 70, 64), ("R", 47, 31, 23),
                                         ("R", 3, 200, 126), ("C", 40, 33, 64), ("C", 7, 5, 3)]):
        ar, ac = (m, k) if lay == "R" else (k, m)
        br, bc = (k, n) if lay == "R" else (n, k)
        cr, cc = (m, n) if lay == "R" else (n, m)
        lda, ldb, ldc = ac + 1, bc + 2, cc + 3
        kind = ["D113", "Dexp"][gi % 2]
        A = qgen.matrix(rng, ar, ac, kind, lda); B = qgen.matrix(rng, br, bc, kind, ldb); C0 = qgen.matrix(rng, cr, cc, kind, ldc)
        alpha, beta = quad.random_quads(rng, 2)
        C = C0.copy()
        ref.gemm(lay, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        p = f"gemm{gi}_"
        out.update({p + "A": A, p + "B": B, p + "C0": C0, p + "C": C, p + "alpha": alpha, p + "beta": beta,
                    p + "dims": np.array([m, n, k, lda, ldb, ldc]), p + "layout": np.array([ord(lay)])})
        cases.append(gi)
    out["gemm_cases"] = np.array(cases)
    # gemv through the reference C ABI (double scalars, transpose relabelling)
    vi = 0
    for lay in "RC":
        for tr in "NT":
            for (m, n, incx, incy) in [(77, 131, 1, 1), (20, 501, 2, 3)]:
                rows, cols = (m, n) if lay == "R" else (n, m)
                lda = cols + 2
                A = qgen.matrix(rng, rows, cols, "D113", lda)
                xn, yn = (n, m) if tr == "N" else (m, n)
                x = quad.random_quads(rng, (xn - 1) * incx + 1); y0 = quad.random_quads(rng, (yn - 1) * incy + 1)
                y = y0.copy()
                ref.c_qgemv(lay, tr, m, n, 1.25, A, lda, x, incx, -0.75, y, incy)
                p = f"gemv{vi}_"
                out.update({p + "A": A, p + "x": x, p + "y0": y0, p + "y": y, p + "dims": np.array([m, n, lda, incx, incy]),
                            p + "lt": np.array([ord(lay), ord(tr)])})
                vi += 1
    out["gemv_count"] = np.array([vi])
    # dot / nrm2 for several thread counts T (the reference's result depends on T)
    di = 0
    for ci, (n, incx, incy) in enumerate([(20011, 1, 1), (499, 1, 1), (1000, 3, 2)]):
        x = quad.random_quads(rng, (n - 1) * incx + 1); y = quad.random_quads(rng, (n - 1) * incy + 1)
        out[f"dotcfg{ci}_x"], out[f"dotcfg{ci}_y"] = x, y
        for T in (1, 2, 3, 8):
            ref.set_num_threads(T)
            p = f"dot{di}_"
            out.update({p + "dims": np.array([n, incx, incy, T, ci]), p + "dot": ref.dot(n, x, incx, y, incy),
                        p + "nrm2": ref.nrm2(n, x, incx), p + "cdot": np.array([ref.c_qdot(n, x, incx, y, incy)])})
            di += 1
    out["dot_count"] = np.array([di])
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
