"""ctypes binding of libqblas_b200.so — the C ABI declared in include/qblas_b200.h.

This is the reference-side binding a Python/numpy FFI consumer would write for the reference's C
entry points (/root/reference/include/quadblas/interface/c_interface.hpp:13 "easy integration with
Python/numpy"), pointed at the CUDA library instead.  There is no fallback: if the shared library
is missing the import of any compute entry point raises.
"""
import ctypes as C
import os

# void (*qb_pass_cb)(int64_t row0, int64_t rows, void *user)  (include/qblas_b200.h)
PASS_CB = C.CFUNCTYPE(None, C.c_int64, C.c_int64, C.c_void_p)
# int (*qb_bpanel_cb)(int64_t col0, int64_t cols, void *stream, const void **panel, int64_t *ld, void *user)
BPANEL_CB = C.CFUNCTYPE(C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqblas_b200.so")


class QbQuad(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64)]


class QblasError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the CDLL; raises QblasError if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QblasError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, ci, cc, cd = C.c_void_p, C.c_int64, C.c_int, C.c_char, C.c_double
    qp = C.POINTER(QbQuad)
    sig = {
        # reference C ABI (c_interface.hpp:21-146)
        "quadblas_qdot": (cd, [ci, vp, ci, vp, ci]),
        "quadblas_qnrm2": (cd, [ci, vp, ci]),
        "quadblas_qaxpy": (None, [ci, cd, vp, ci, vp, ci]),
        "quadblas_qgemv": (None, [cc, cc, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]),
        "quadblas_qgemm": (None, [cc, cc, cc, ci, ci, ci, cd, vp, ci, vp, ci, cd, vp, ci]),
        "quadblas_set_num_threads": (None, [ci]),
        "quadblas_get_num_threads": (ci, []),
        "quadblas_get_version": (C.c_char_p, []),
        "quadblas_is_aligned": (ci, [vp]),
        # extended API
        "qb_init": (ci, []),
        "qb_last_error": (C.c_char_p, []),
        "qb_last_error_code": (ci, []),
        "qb_clear_error": (None, []),
        "qb_build_info": (C.c_char_p, []),
        "qb_set_mode": (None, [ci]),
        "qb_get_mode": (ci, []),
        "qb_set_kc": (None, [ci]),
        "qb_get_kc": (ci, []),
        "qb_set_honor_trans": (None, [ci]),
        "qb_get_honor_trans": (ci, []),
        "qb_set_tensor_path": (None, [ci]),
        "qb_get_tensor_path": (ci, []),
        "qb_set_gemm_pass_callback": (None, [PASS_CB, vp, ci]),
        "qb_set_gemm_b_panels": (None, [BPANEL_CB, vp, i64, vp]),
        "qb_gemm_colstats_dev": (ci, [cc, cc, i64, i64, vp, i64, vp, vp]),
        "qb_crt_plan": (ci, [ci, ci, i64, C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]),
        "qb_crt_residues_dev": (ci, [cc, cc, i64, i64, vp, i64, vp, ci, ci, vp, i64, vp]),
        "qb_set_gemm_b_planes": (None, [ci, ci]),
        "qb_set_tensor_window": (None, [ci]),
        "qb_get_tensor_window": (ci, []),
        "qb_set_tensor_unit": (None, [i64, i64]),
        "qb_get_tensor_unit": (None, [C.POINTER(i64), C.POINTER(i64)]),
        "qb_set_tensor_ramp": (None, [i64, i64]),
        "qb_get_tensor_ramp": (None, [C.POINTER(i64), C.POINTER(i64)]),
        "qb_set_tensor_workspace_limit": (None, [C.c_size_t]),
        "qb_get_tensor_workspace_limit": (C.c_size_t, []),
        "qb_dot_kernel": (ci, [i64, vp, vp, qp]),
        "qb_peer_alloc": (vp, [C.c_size_t]),
        "qb_peer_free": (None, [vp]),
        "qb_peer_export": (ci, [vp, vp]),
        "qb_peer_open": (vp, [vp]),
        "qb_peer_close": (ci, [vp]),
        "qb_set_gemm_peer_outputs": (ci, [ci, C.POINTER(vp)]),
        "qb_get_gemm_peer_written": (ci, []),
        "qb_set_host_slabs": (None, [ci]),
        "qb_get_host_slabs": (ci, []),
        "qb_set_tensor_pass_shape": (None, [ci]),
        "qb_get_tensor_pass_shape": (ci, []),
        "qb_crt_pass_rows": (ci, [i64, i64, ci, C.POINTER(i64), ci]),
        "qb_set_fast_variant": (None, [ci]),
        "qb_get_fast_variant": (ci, []),
        "qb_set_ref_gemm_kernel": (None, [ci]),
        "qb_get_ref_gemm_kernel": (ci, []),
        "qb_gemv_last_declined": (C.c_int64, []),
        "qb_set_beta0_classes": (None, [ci]),
        "qb_get_beta0_classes": (ci, []),
        "qb_oz_last_stats": (None, [C.POINTER(i64)]),
        "qb_oz_last_mma_ms": (cd, [C.POINTER(ci)]),
        "qb_oz_last_mma_timeline": (ci, [C.POINTER(cd), ci]),
        "qb_oz_i8gemm_dev": (ci, [vp, vp, ci, ci, i64, i64, i64, i64, i64, vp, i64, i64, vp]),
        "qb_gemm": (ci, [cc, cc, cc, i64, i64, i64, qp, vp, i64, vp, i64, qp, vp, i64]),
        "qb_gemv": (ci, [cc, i64, i64, qp, vp, i64, vp, i64, qp, vp, i64]),
        "qb_dot": (ci, [i64, vp, i64, vp, i64, qp]),
        "qb_nrm2": (ci, [i64, vp, i64, qp]),
        "qb_axpy": (ci, [i64, qp, vp, i64, vp, i64]),
        "qb_gemm_dev": (ci, [cc, cc, cc, i64, i64, i64, qp, vp, i64, vp, i64, qp, vp, i64, vp]),
        "qb_gemv_dev": (ci, [cc, i64, i64, qp, vp, i64, vp, i64, qp, vp, i64, vp]),
        "qb_gemv_rows_dev": (ci, [cc, i64, i64, qp, vp, i64, vp, i64, qp, vp, i64, vp, i64]),
        "qb_dot_dev": (ci, [i64, vp, i64, vp, i64, vp, vp]),
        "qb_nrm2_dev": (ci, [i64, vp, i64, vp, vp]),
        "qb_axpy_dev": (ci, [i64, qp, vp, i64, vp, i64, vp]),
        "qb_dot_partials_dev": (ci, [i64, vp, i64, vp, i64, i64, i64, vp, vp]),
        "qb_fold_partials_dev": (ci, [i64, vp, ci, vp, vp]),
        "qb_elementwise_dev": (ci, [ci, i64, vp, vp, vp, vp, vp]),
        "qb_host_alloc": (vp, [C.c_size_t]),
        "qb_host_free": (None, [vp]),
        "qb_from_double": (QbQuad, [cd]),
        "qb_to_double": (cd, [QbQuad]),
        "qb_launch_count": (i64, []),
        "qb_fma_microbench_dev": (ci, [ci, ci, ci, ci, vp, C.POINTER(i64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._qb_signatures = sig
    _lib = L
    return L


def check(rc, what="qblas_b200 call"):
    if rc != 0:
        L = lib()
        msg = L.qb_last_error().decode()
        L.qb_clear_error()      # the error is reported here: it must not make a later, successful call look failed
        raise QblasError(f"{what} failed (code {rc}): {msg}")
