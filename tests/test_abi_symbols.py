"""The C-ABI library loads on a CPU-only box and exports every symbol include/qblas_b200.h
declares; non-compute entry points behave like the reference's; compute entry points fail loudly
(never fall back to a CPU path) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import qblas_b200
from qblas_b200 import quad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qblas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:qb|quadblas)_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = qblas_b200.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"libqblas_b200.so does not export {n}"
    assert set(names) == set(L._qb_signatures), "ctypes binding and header disagree"


def test_version_threads_alignment():
    assert qblas_b200.quadblas_get_version() == "QuadBLAS 1.0.0 - High Performance Quad Precision BLAS"  # c_interface.hpp:136
    # default = omp_get_max_threads() as the reference reads it (threading/openmp_utils.hpp:10-17): OMP_NUM_THREADS, else the hardware
    env_t = os.environ.get("OMP_NUM_THREADS", "").split(",")[0].strip()
    assert qblas_b200.quadblas_get_num_threads() == (int(env_t) if env_t.isdigit() and int(env_t) > 0 else (os.cpu_count() or 1))
    qblas_b200.quadblas_set_num_threads(5)
    assert qblas_b200.quadblas_get_num_threads() == 5
    qblas_b200.quadblas_set_num_threads(0)
    L = qblas_b200.lib()
    assert L.quadblas_is_aligned(C.c_void_p(64)) == 1 and L.quadblas_is_aligned(C.c_void_p(48)) == 0  # ALIGNMENT = 32


def test_mode_and_kc_setters():
    qblas_b200.set_mode(qblas_b200.MODE_FAST); assert qblas_b200.get_mode() == 1
    qblas_b200.set_mode(qblas_b200.MODE_REFERENCE); assert qblas_b200.get_mode() == 0
    L = qblas_b200.lib()
    assert L.qb_get_kc() == 126
    L.qb_set_kc(256); assert L.qb_get_kc() == 256
    L.qb_set_kc(126)
    assert L.qb_get_honor_trans() == 0


def test_casts_match_oracle(oracle):
    L = qblas_b200.lib()
    rng = np.random.default_rng(1)
    vals = list(rng.standard_normal(200)) + [0.0, -0.0, 5e-324, -2.2250738585072014e-308, 1.7976931348623157e308, float("inf")]
    for d in vals:
        q = L.qb_from_double(float(d))
        o = oracle.from_double(d)
        assert (q.lo, q.hi) == (int(o[0]), int(o[1]))
        assert L.qb_to_double(q) == d or (d != d)
    # narrowing with rounding: random full-mantissa quads, incl. the double-subnormal range
    qs = quad.random_quads(rng, 4000, "D113", -1100, 1030)
    for v in qs:
        got = L.qb_to_double(qblas_b200._lib.QbQuad(lo=int(v[0]), hi=int(v[1])))
        assert got == oracle.to_double(v)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = quad.from_double(np.ones(4))
    with pytest.raises(qblas_b200.QblasError):
        qblas_b200.dot(4, x, 1, x, 1)
    with pytest.raises(qblas_b200.QblasError):
        qblas_b200.quadblas_qdot(4, x, 1, x, 1)
    y = x.copy()
    with pytest.raises(qblas_b200.QblasError):
        qblas_b200.quadblas_qgemv("R", "N", 2, 2, 1.0, x, 2, x, 1, 0.0, y, 1)
    assert quad.same_bits(x, y).all()  # outputs untouched on failure


def test_environment_selects_mode_kc_and_threads():
    """An UNMODIFIED caller of the reference API picks the numerical mode through the environment (QUADBLAS_MODE, QUADBLAS_KC), and
    the default thread count follows OMP_NUM_THREADS like omp_get_max_threads() does in the reference."""
    import subprocess
    import sys
    code = ("import qblas_b200 as q; L = q.lib(); print(L.qb_get_mode(), L.qb_get_kc(), L.quadblas_get_num_threads()); "
            "L.quadblas_set_num_threads(3); print(L.quadblas_get_num_threads())")
    def run(env):
        e = dict(os.environ, PYTHONPATH=ROOT); e.pop("OMP_NUM_THREADS", None); e.pop("QUADBLAS_MODE", None); e.pop("QUADBLAS_KC", None)
        e.update(env)
        out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, check=True).stdout.split()
        return [int(v) for v in out]
    assert run({}) == [0, 126, os.cpu_count() or 1, 3]
    assert run({"QUADBLAS_MODE": "fast", "QUADBLAS_KC": "256", "OMP_NUM_THREADS": "7"}) == [1, 256, 7, 3]
    assert run({"QUADBLAS_MODE": "REFERENCE", "OMP_NUM_THREADS": "4,2"}) == [0, 126, 4, 3]
    assert run({"QUADBLAS_MODE": "bogus", "QUADBLAS_KC": "-5", "OMP_NUM_THREADS": "zero"}) == [0, 126, os.cpu_count() or 1, 3]


def test_a_reported_error_does_not_stick():
    """A failed call raises once; the thread-local error is cleared with the report, and the reference-named entry points clear it at
    entry, so qb_last_error_code() always describes the LAST call only."""
    import torch
    L = qblas_b200.lib()
    x = quad.from_double(np.ones(4))
    with pytest.raises(qblas_b200.QblasError):
        qblas_b200.gemv("R", -1, 2, 1.0, x, 2, x, 1, 0.0, x.copy(), 1)        # QB_ERR_ARG, no GPU needed
    assert L.qb_last_error_code() == 0 and L.qb_last_error() == b""
    L.qb_gemv(b"R", -1, 2, None, None, 2, None, 1, None, None, 1)              # a C caller that ignores the return code
    assert L.qb_last_error_code() == 2
    if not torch.cuda.is_available():
        L.quadblas_qaxpy(0, 1.0, None, 1, None, 1)                             # n <= 0: returns at once (c_interface.hpp:49) ...
        assert L.qb_last_error_code() == 0                                     # ... after clearing the stale code
    L.qb_clear_error()


def test_residue_scheme_row_pass_partition():
    """Host logic of the row passes (csrc/qb_ozaki.cu, oz_crt_pass_rows): shape 0 is the equal split the kernels were measured with,
    shape 1 (experimental) shortens the first and the last pass; every partition covers the rows exactly once in tile multiples."""
    from qblas_b200 import api
    assert api.get_tensor_pass_shape() == 0
    for m in (1, 127, 128, 384, 1000, 4096, 8192, 8200, 32768, 65536 + 5):
        for cap in (128, 256, 1024, 2048, 8192):
            for shape in (0, 1):
                rows = api.crt_pass_rows(m, cap, shape)
                assert sum(rows) == m and all(0 < r <= cap for r in rows), (m, cap, shape, rows)
                if shape == 0:      # the old loop: r0 += cap, mr = min(cap, m - r0)
                    assert rows == [min(cap, m - r0) for r0 in range(0, m, cap)]
                # a pass that is not the last in ROW ORDER starts the next one on a tile boundary
                r0 = 0
                for r in rows[:-1]:
                    r0 += r
                    assert r0 % 128 == 0, (m, cap, shape, rows)
    assert api.crt_pass_rows(8192, 2048, 1) == [512, 2048, 2048, 2048, 1024, 512]
    assert api.crt_pass_rows(384, 128, 1) == [128, 128, 128]          # too small to shape


def test_measured_defaults():
    """The settings every number under profiles/ was measured with are the library defaults."""
    L = qblas_b200.lib()
    assert L.qb_get_mode() == 0                   # reference order (bit exact) unless fast mode is requested
    assert L.qb_get_tensor_path() == 1            # fast-mode qgemm: tensor path for m, n >= 128, k >= 256
    assert L.qb_get_tensor_window() == 144        # bits per operand window when the spans do not fit the moduli
    assert qblas_b200.get_tensor_unit() == (2048, 4096) and qblas_b200.get_tensor_ramp() == (0, 0)   # pipeline unit: A pass x B panel
    assert L.qb_get_tensor_pass_shape() == 0      # equal row passes
    assert L.qb_get_fast_variant() == 2           # large row-major qgemv: sliced FP64 accumulate; qdot / qnrm2 / the rest: window accumulator
    assert L.qb_get_ref_gemm_kernel() == 1        # reference-order qgemm: the branch-free kernel (k_gemm_nb); 0 = the first version
    assert L.qb_get_gemm_peer_written() == 0
    assert L.qb_get_host_slabs() == 8             # pipelined all-host qgemm / qgemv: eight slabs
