"""Golden vectors produced by the reference itself (tests/golden/make_golden.py) — checked on the
oracle here (CPU) and on the CUDA path in the -m gpu twin below."""
import os

import numpy as np
import pytest

from qblas_b200 import quad

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
REGS = ["any", "similar", "cancel", "gaps", "tiny", "subres", "overflow", "nearovf", "specials"]


def _gemm_case(gi):
    p = f"gemm{gi}_"
    m, n, k, lda, ldb, ldc = (int(v) for v in G[p + "dims"])
    return chr(int(G[p + "layout"][0])), m, n, k, lda, ldb, ldc, G[p + "alpha"], G[p + "beta"], G[p + "A"], G[p + "B"], G[p + "C0"], G[p + "C"]


def test_oracle_vs_golden(oracle):
    for reg in REGS:
        assert quad.same_bits(oracle.fma(G[f"fma_{reg}_a"], G[f"fma_{reg}_b"], G[f"fma_{reg}_c"]), G[f"fma_{reg}_out"]).all(), reg
    for gi in G["gemm_cases"]:
        lay, m, n, k, lda, ldb, ldc, alpha, beta, A, B, C0, C = _gemm_case(int(gi))
        Co = C0.copy(); oracle.gemm(lay, m, n, k, alpha, A, lda, B, ldb, beta, Co, ldc)
        assert quad.same_bits(Co, C).all(), gi
    for vi in range(int(G["gemv_count"][0])):
        p = f"gemv{vi}_"
        m, n, lda, incx, incy = (int(v) for v in G[p + "dims"]); lay, tr = (chr(int(v)) for v in G[p + "lt"])
        y = G[p + "y0"].copy(); oracle.c_qgemv(lay, tr, m, n, 1.25, G[p + "A"], lda, G[p + "x"], incx, -0.75, y, incy)
        assert quad.same_bits(y, G[p + "y"]).all(), vi
    for di in range(int(G["dot_count"][0])):
        p = f"dot{di}_"
        n, incx, incy, T, ci = (int(v) for v in G[p + "dims"])
        x, y = G[f"dotcfg{ci}_x"], G[f"dotcfg{ci}_y"]
        assert quad.same_bits(oracle.dot(n, x, incx, y, incy, T), G[p + "dot"]).all()
        assert quad.same_bits(oracle.nrm2(n, x, incx, T), G[p + "nrm2"]).all()
        assert oracle.to_double(G[p + "dot"]) == float(G[p + "cdot"][0])


@pytest.mark.gpu
def test_cuda_vs_golden(qb):
    import torch
    from gpu_util import to_dev, to_host
    for reg in REGS:
        a = G[f"fma_{reg}_a"]
        out = torch.empty((a.shape[0], 2), dtype=torch.int64, device="cuda")
        for op in (0, 1):
            qb.elementwise(op, to_dev(a), to_dev(G[f"fma_{reg}_b"]), to_dev(G[f"fma_{reg}_c"]), out)
            assert quad.same_bits(to_host(out), G[f"fma_{reg}_out"]).all(), (reg, op)
    qb.set_mode(qb.MODE_REFERENCE)
    for gi in G["gemm_cases"]:
        lay, m, n, k, lda, ldb, ldc, alpha, beta, A, B, C0, C = _gemm_case(int(gi))
        Cg = C0.copy(); qb.gemm(lay, m, n, k, alpha, A, lda, B, ldb, beta, Cg, ldc)
        assert quad.same_bits(Cg, C).all(), gi
    for vi in range(int(G["gemv_count"][0])):
        p = f"gemv{vi}_"
        m, n, lda, incx, incy = (int(v) for v in G[p + "dims"]); lay, tr = (chr(int(v)) for v in G[p + "lt"])
        y = G[p + "y0"].copy(); qb.quadblas_qgemv(lay, tr, m, n, 1.25, G[p + "A"].copy(), lda, G[p + "x"].copy(), incx, -0.75, y, incy)
        assert quad.same_bits(y, G[p + "y"]).all(), vi
    try:
        for di in range(int(G["dot_count"][0])):
            p = f"dot{di}_"
            n, incx, incy, T, ci = (int(v) for v in G[p + "dims"])
            qb.quadblas_set_num_threads(T)
            x, y = G[f"dotcfg{ci}_x"].copy(), G[f"dotcfg{ci}_y"].copy()
            assert quad.same_bits(qb.dot(n, x, incx, y, incy), G[p + "dot"]).all()
            assert quad.same_bits(qb.nrm2(n, x, incx), G[p + "nrm2"]).all()
            assert qb.quadblas_qdot(n, x, incx, y, incy) == float(G[p + "cdot"][0])
    finally:
        qb.quadblas_set_num_threads(0)
