"""Python mirror of the reference's operator interface for the binary128 hot path.

Same names, argument meaning and marshalling as the reference's C entry points
(/root/reference/include/quadblas/interface/c_interface.hpp:21-146) plus quad-typed variants that
mirror QuadBLAS::gemm/gemv/dot/axpy and Vector::dot/norm (cpp_classes.hpp:66-81,143-154).

Operands are binary128 bit patterns:
  * numpy uint64 arrays of shape (..., 2)  -> host path (the library stages H2D/D2H itself), or
  * torch CUDA tensors (int64, shape (..., 2)) -> device path, asynchronous on torch's current
    stream, no host synchronisation.
Every call goes through libqblas_b200.so; nothing here computes.
"""
import ctypes as C

import numpy as np

from . import quad
from ._lib import QbQuad, QblasError, check, lib

MODE_REFERENCE = 0
MODE_FAST = 1


# ------------------------------------------------------------------ scalars / pointers
def _q(v):
    """Python float | (hi, lo) tuple | numpy (2,) uint64 -> QbQuad (exact)."""
    if isinstance(v, QbQuad):
        return v
    if isinstance(v, (float, int)) and not isinstance(v, bool):
        return lib().qb_from_double(float(v))
    if isinstance(v, tuple):
        return QbQuad(lo=int(v[1]), hi=int(v[0]))
    a = np.asarray(v, dtype=np.uint64).reshape(2)
    return QbQuad(lo=int(a[0]), hi=int(a[1]))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if _is_torch(x):
        if not x.is_cuda:
            raise TypeError("torch operands must be CUDA tensors (use numpy arrays for host data)")
        if x.element_size() * x.shape[-1] != 16:
            raise TypeError("torch quad tensors must have a trailing dimension of 16 bytes (int64 x 2)")
        return C.c_void_p(x.data_ptr())
    if not (isinstance(x, np.ndarray) and x.dtype == np.uint64 and x.shape[-1] == 2):
        raise TypeError("host quads must be numpy uint64 arrays of shape (..., 2)")
    return C.c_void_p(x.ctypes.data)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c(ch):
    return ch.encode() if isinstance(ch, str) else ch


# ------------------------------------------------------------------ configuration
def init():
    check(lib().qb_init(), "qb_init")


def set_mode(mode):
    lib().qb_set_mode(int(mode))


def get_mode():
    return lib().qb_get_mode()


def set_kc(kc):
    lib().qb_set_kc(int(kc))


def set_honor_trans(on):
    lib().qb_set_honor_trans(1 if on else 0)


def quadblas_set_num_threads(t):
    lib().quadblas_set_num_threads(int(t))


def quadblas_get_num_threads():
    return lib().quadblas_get_num_threads()


def quadblas_get_version():
    return lib().quadblas_get_version().decode()


def quadblas_is_aligned(arr):
    return lib().quadblas_is_aligned(_ptr(arr))


def launch_count():
    return lib().qb_launch_count()


# ------------------------------------------------------------------ reference C ABI (double scalars)
def quadblas_qdot(n, x, incx, y, incy):
    r = lib().quadblas_qdot(n, _ptr(x), incx, _ptr(y), incy)
    _raise_if_error("quadblas_qdot")
    return r


def quadblas_qnrm2(n, x, incx):
    r = lib().quadblas_qnrm2(n, _ptr(x), incx)
    _raise_if_error("quadblas_qnrm2")
    return r


def quadblas_qaxpy(n, alpha, x, incx, y, incy):
    lib().quadblas_qaxpy(n, float(alpha), _ptr(x), incx, _ptr(y), incy)
    _raise_if_error("quadblas_qaxpy")


def quadblas_qgemv(layout, trans, m, n, alpha, A, lda, x, incx, beta, y, incy):
    lib().quadblas_qgemv(_c(layout), _c(trans), m, n, float(alpha), _ptr(A), lda, _ptr(x), incx, float(beta), _ptr(y), incy)
    _raise_if_error("quadblas_qgemv")


def quadblas_qgemm(layout, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C_, ldc):
    lib().quadblas_qgemm(_c(layout), _c(transa), _c(transb), m, n, k, float(alpha), _ptr(A), lda, _ptr(B), ldb,
                         float(beta), _ptr(C_), ldc)
    _raise_if_error("quadblas_qgemm")


def _raise_if_error(what):
    L = lib()
    code = L.qb_last_error_code()
    if code:
        msg = L.qb_last_error().decode()
        L.qb_clear_error()
        from ._lib import QblasError
        raise QblasError(f"{what} failed (code {code}): {msg}")


# ------------------------------------------------------------------ quad-typed API (QuadBLAS:: free functions)
def gemm(layout, m, n, k, alpha, A, lda, B, ldb, beta, C_, ldc, transa="N", transb="N"):
    """QuadBLAS::gemm (level3.hpp:215).  Device tensors -> async on the current torch stream."""
    L = lib()
    a, b = _q(alpha), _q(beta)
    if _is_torch(C_):
        check(L.qb_gemm_dev(_c(layout), _c(transa), _c(transb), m, n, k, C.byref(a), _ptr(A), lda, _ptr(B), ldb,
                            C.byref(b), _ptr(C_), ldc, _stream()), "qb_gemm_dev")
    else:
        check(L.qb_gemm(_c(layout), _c(transa), _c(transb), m, n, k, C.byref(a), _ptr(A), lda, _ptr(B), ldb,
                        C.byref(b), _ptr(C_), ldc), "qb_gemm")


def gemv(layout, m, n, alpha, A, lda, x, incx, beta, y, incy, m_total=0):
    """QuadBLAS::gemv (level2.hpp:85) — AFTER any transpose relabelling.  m_total (device tensors): these m rows are a block of a
    qgemv with m_total rows; the fast-mode kernel is planned for the whole, so the block carries the bits of the unsplit call."""
    L = lib()
    a, b = _q(alpha), _q(beta)
    if _is_torch(y):
        if m_total:
            check(L.qb_gemv_rows_dev(_c(layout), m, n, C.byref(a), _ptr(A), lda, _ptr(x), incx, C.byref(b), _ptr(y), incy,
                                     _stream(), int(m_total)), "qb_gemv_rows_dev")
            return
        check(L.qb_gemv_dev(_c(layout), m, n, C.byref(a), _ptr(A), lda, _ptr(x), incx, C.byref(b), _ptr(y), incy,
                            _stream()), "qb_gemv_dev")
    else:
        check(L.qb_gemv(_c(layout), m, n, C.byref(a), _ptr(A), lda, _ptr(x), incx, C.byref(b), _ptr(y), incy), "qb_gemv")


def dot(n, x, incx, y, incy, out=None):
    """QuadBLAS::dot / Vector::dot (level1.hpp:80, cpp_classes.hpp:66): full binary128 result.
    Host arrays -> returns numpy (2,) uint64.  Device tensors -> writes `out` (device, 16 B), async."""
    L = lib()
    if _is_torch(x):
        check(L.qb_dot_dev(n, _ptr(x), incx, _ptr(y), incy, _ptr(out), _stream()), "qb_dot_dev")
        return out
    r = QbQuad()
    check(L.qb_dot(n, _ptr(x), incx, _ptr(y), incy, C.byref(r)), "qb_dot")
    return np.array([r.lo, r.hi], dtype=np.uint64)


def nrm2(n, x, incx, out=None):
    """Vector::norm (cpp_classes.hpp:78): sqrt(dot(x,x)), full binary128 result."""
    L = lib()
    if _is_torch(x):
        check(L.qb_nrm2_dev(n, _ptr(x), incx, _ptr(out), _stream()), "qb_nrm2_dev")
        return out
    r = QbQuad()
    check(L.qb_nrm2(n, _ptr(x), incx, C.byref(r)), "qb_nrm2")
    return np.array([r.lo, r.hi], dtype=np.uint64)


def axpy(n, alpha, x, incx, y, incy):
    """QuadBLAS::axpy (level1.hpp:190)."""
    L = lib()
    a = _q(alpha)
    if _is_torch(y):
        check(L.qb_axpy_dev(n, C.byref(a), _ptr(x), incx, _ptr(y), incy, _stream()), "qb_axpy_dev")
    else:
        check(L.qb_axpy(n, C.byref(a), _ptr(x), incx, _ptr(y), incy), "qb_axpy")


def dot_partials(n_local, x, incx, y, incy, chunk, nchunks, out):
    """Reference-order chunk partials of a local shard (see qb_dot_partials_dev)."""
    check(lib().qb_dot_partials_dev(n_local, _ptr(x), incx, _ptr(y), incy, chunk, nchunks, _ptr(out), _stream()), "qb_dot_partials_dev")
    return out


def fold_partials(count, partials, out, do_sqrt=False):
    """Exchange step of a sharded dot: fold `count` device partials in index order."""
    check(lib().qb_fold_partials_dev(count, _ptr(partials), 1 if do_sqrt else 0, _ptr(out), _stream()), "qb_fold_partials_dev")
    return out


def elementwise(op, a, b, c, out):
    """Scalar-op probe on device tensors: op 0 fma, 1 chain fma, 2 mul, 3 add, 4 sqrt, 5 cast round trip."""
    n = out.numel() // 2
    pb = _ptr(b) if b is not None else C.c_void_p(0)
    pc = _ptr(c) if c is not None else C.c_void_p(0)
    check(lib().qb_elementwise_dev(op, n, _ptr(a), pb, pc, _ptr(out), _stream()), "qb_elementwise_dev")
    return out


def fma_microbench(variant, blocks, threads, iters, sink):
    n = C.c_int64(0)
    check(lib().qb_fma_microbench_dev(variant, blocks, threads, iters, _ptr(sink), C.byref(n), _stream()), "qb_fma_microbench_dev")
    return n.value


# ------------------------------------------------------------------ fast-mode tensor path (csrc/qb_ozaki.cu)
TENSOR_OFF, TENSOR_AUTO, TENSOR_ALWAYS = 0, 1, 2


def set_tensor_path(v):
    lib().qb_set_tensor_path(int(v))


def get_tensor_path():
    return lib().qb_get_tensor_path()


_pass_cb_ref = None


def set_gemm_pass_callback(fn, min_passes=1):
    """Row-pass hook of the device qgemm (qb_set_gemm_pass_callback): fn(row0, rows) is called after the work of each row pass has been
    enqueued; fn = None removes it.  The ctypes thunk is kept alive here."""
    global _pass_cb_ref
    from ._lib import PASS_CB
    if fn is None:
        lib().qb_set_gemm_pass_callback(C.cast(None, PASS_CB), None, 1)
        _pass_cb_ref = None
        return
    thunk = PASS_CB(lambda r0, rows, user: fn(int(r0), int(rows)))
    lib().qb_set_gemm_pass_callback(thunk, None, int(min_passes))
    _pass_cb_ref = thunk


_bpanel_cb_ref = None


def set_gemm_b_panels(fn, panel_cols=0, colstats=None):
    """Streamed B of the device qgemm (qb_set_gemm_b_panels): fn(col0, cols, stream_ptr) -> (panel_ptr, ld) is called before the library
    reads columns [col0, col0 + cols) of op(B); it must make the CUDA stream `stream_ptr` wait for the panel's arrival (e.g.
    torch.cuda.ExternalStream(stream_ptr).wait_event(...)).  colstats: device int32 tensor of 3 n entries from gemm_colstats.
    fn = None removes the hook.  The ctypes thunk is kept alive here."""
    global _bpanel_cb_ref
    from ._lib import BPANEL_CB
    if fn is None:
        lib().qb_set_gemm_b_panels(C.cast(None, BPANEL_CB), None, 0, None)
        _bpanel_cb_ref = None
        return

    def thunk(col0, cols, stream, panel, ld, user):
        try:
            ptr, ldv = fn(int(col0), int(cols), int(stream or 0))
            panel[0] = int(ptr); ld[0] = int(ldv)
            return 0
        except Exception as e:   # never let an exception cross the C ABI
            import sys
            print(f"[qblas_b200] B-panel callback failed: {e!r}", file=sys.stderr, flush=True)
            return 1

    cb = BPANEL_CB(thunk)
    lib().qb_set_gemm_b_panels(cb, None, int(panel_cols), C.c_void_p(colstats.data_ptr()))
    _bpanel_cb_ref = (cb, colstats)


def gemm_colstats(layout, k, n, B, ldb, out, transb="N"):
    """Column statistics of op(B) (k x n) for set_gemm_b_panels: `out` = device int32 tensor with 3 n entries."""
    check(lib().qb_gemm_colstats_dev(_c(layout), _c(transb), k, n, _ptr(B), ldb, C.c_void_p(out.data_ptr()), _stream()), "qb_gemm_colstats_dev")
    return out


def crt_plan(span_a, span_b, k):
    """(moduli, window_a, window_b, truncated mask) the tensor path's planner gives for these bit spans (qb_crt_plan); None: no plan."""
    n, wa, wb = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().qb_crt_plan(int(span_a), int(span_b), int(k), C.byref(n), C.byref(wa), C.byref(wb))
    return None if rc < 0 else (int(n.value), int(wa.value), int(wb.value), int(rc))


def crt_residues(layout, k, n, B, ldb, emax, window, moduli, planes, plane_stride=0, transb="N", data_ptr=None):
    """Residue planes of the n columns of op(B) (qb_crt_residues_dev).  B: device quads (or data_ptr = a raw device address);
    emax: int32 device tensor (first array of the columns' statistics); planes: int8 device tensor or raw address."""
    bp = C.c_void_p(int(data_ptr)) if data_ptr is not None else _ptr(B)
    pp = C.c_void_p(int(planes)) if isinstance(planes, int) else C.c_void_p(planes.data_ptr())
    check(lib().qb_crt_residues_dev(_c(layout), _c(transb), k, n, bp, ldb, C.c_void_p(emax.data_ptr()), int(window), int(moduli), pp, int(plane_stride), _stream()),
          "qb_crt_residues_dev")


def set_gemm_b_planes(moduli=0, window=0):
    """The panels of set_gemm_b_panels are residue planes computed with `window` and `moduli` moduli (0: element panels)."""
    lib().qb_set_gemm_b_planes(int(moduli), int(window))


def set_tensor_window(bits):
    """Bits per operand window the tensor-path planner grants when the operands' spans do not fit the moduli (default 144)."""
    lib().qb_set_tensor_window(int(bits))


def get_tensor_window():
    return lib().qb_get_tensor_window()


def set_tensor_unit(rows=0, cols=0):
    """Pipeline unit of the tensor path: rows of an A pass x columns of a B panel (default 2048 x 4096)."""
    lib().qb_set_tensor_unit(int(rows), int(cols))


def get_tensor_unit():
    r, c = C.c_int64(0), C.c_int64(0)
    lib().qb_get_tensor_unit(C.byref(r), C.byref(c))
    return int(r.value), int(c.value)


def set_tensor_ramp(rows=0, cols=0):
    """Rows of the first pass / columns of the first panel of the tensor path (0 = like the others)."""
    lib().qb_set_tensor_ramp(int(rows), int(cols))


def get_tensor_ramp():
    r, c = C.c_int64(0), C.c_int64(0)
    lib().qb_get_tensor_ramp(C.byref(r), C.byref(c))
    return int(r.value), int(c.value)


def set_tensor_workspace_limit(nbytes=0):
    """Cap of the tensor path's workspace in bytes (0 = 85 % of the free device memory)."""
    lib().qb_set_tensor_workspace_limit(int(nbytes))


def dot_kernel(n, x, y):
    """QuadBLAS::dot_kernel_vectorized (level1.hpp:14-35) on host arrays: the two-lane reference kernel whatever the mode / T."""
    r = QbQuad()
    check(lib().qb_dot_kernel(n, _ptr(x), _ptr(y), C.byref(r)), "qb_dot_kernel")
    return np.array([r.lo, r.hi], dtype=np.uint64)


def set_gemm_peer_outputs(ptrs):
    """Fused gather (qb_set_gemm_peer_outputs): ptrs[q] = address inside peer q's mapped buffer that corresponds to the C argument of the
    following device qgemm calls; None / [] switches it off."""
    ptrs = list(ptrs or [])
    arr = (C.c_void_p * max(1, len(ptrs)))(*[C.c_void_p(int(p)) for p in ptrs])
    check(lib().qb_set_gemm_peer_outputs(len(ptrs), arr), "qb_set_gemm_peer_outputs")


def gemm_peer_written():
    """Peers written by the last device qgemm (0: it ran a path without the fused stores)."""
    return lib().qb_get_gemm_peer_written()


class _CudaArray:
    """__cuda_array_interface__ carrier so torch can view library-allocated device memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def tensor_from_ptr(ptr, nbytes):
    """torch uint8 view of `nbytes` of device memory at a raw address (e.g. a multicast mapping); the caller keeps the memory alive."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))


def peer_alloc(nbytes):
    """(ptr, torch uint8 view) of exportable device memory (qb_peer_alloc); free with peer_free(ptr) after the views are gone."""
    import torch
    p = lib().qb_peer_alloc(int(nbytes))
    if not p:
        raise QblasError(f"qb_peer_alloc({nbytes}): {lib().qb_last_error().decode()}")
    return int(p), torch.as_tensor(_CudaArray(p, nbytes), device=torch.device("cuda", torch.cuda.current_device()))


def peer_free(ptr):
    lib().qb_peer_free(C.c_void_p(int(ptr)))


def peer_export(ptr):
    h = (C.c_ubyte * 64)()
    check(lib().qb_peer_export(C.c_void_p(int(ptr)), h), "qb_peer_export")
    return bytes(h)


def peer_open(handle):
    buf = (C.c_ubyte * 64)(*handle)
    p = lib().qb_peer_open(buf)
    if not p:
        raise QblasError(f"qb_peer_open: {lib().qb_last_error().decode()}")
    return int(p)


def peer_close(ptr):
    check(lib().qb_peer_close(C.c_void_p(int(ptr))), "qb_peer_close")


def set_host_slabs(slabs):
    """C slabs of the pipelined all-host qgemm (default 4)."""
    lib().qb_set_host_slabs(int(slabs))


def get_host_slabs():
    return lib().qb_get_host_slabs()


def set_tensor_pass_shape(shape):
    """Residue scheme row passes: 0 = equal (default, measured), 1 = short first / last pass (experimental)."""
    lib().qb_set_tensor_pass_shape(int(shape))


def get_tensor_pass_shape():
    return lib().qb_get_tensor_pass_shape()


def crt_pass_rows(m, cap, shape=0):
    """The row-pass partition the residue scheme uses for m rows with at most cap rows per pass."""
    out = (C.c_int64 * 4096)()
    cnt = lib().qb_crt_pass_rows(int(m), int(cap), int(shape), out, 4096)
    return [int(out[i]) for i in range(min(cnt, 4096))]


def set_fast_variant(v):
    """Fast-mode accumulate of dot/nrm2/gemv: 2 (default) = sliced FP64 accumulate for large row-major gemv (csrc/qslice.cuh), window
    accumulator elsewhere; 1 = window accumulator everywhere (csrc/qwide.cuh); 0 = rounded-FMA chains."""
    lib().qb_set_fast_variant(int(v))


def get_fast_variant():
    return lib().qb_get_fast_variant()


def set_ref_gemm_kernel(v):
    """Kernel of the reference-order qgemm (same bits): 1 (default) = k_gemm_nb (branch-free step), 0 = k_gemm (first version)."""
    lib().qb_set_ref_gemm_kernel(int(v))


def get_ref_gemm_kernel():
    return lib().qb_get_ref_gemm_kernel()


def set_beta0_classes(v):
    """Pipelined all-host qgemm with beta = 0: 1 (default) = C_in goes up as one class byte per element, 0 = as its 16 bytes."""
    lib().qb_set_beta0_classes(int(v))


def get_beta0_classes():
    return lib().qb_get_beta0_classes()


def gemv_last_declined():
    """Rows of the last device qgemv that the sliced FP64 kernel declined (recomputed by the window kernel); -1 if that call did not
    take the sliced path.  Synchronises the device."""
    return int(lib().qb_gemv_last_declined())


def oz_last_stats():
    """Plan of the last tensor-path qgemm (qb_oz_last_stats).  pairs = moduli = int8 GEMMs per product (0: the tensor path did not run)."""
    out = (C.c_int64 * 16)()
    lib().qb_oz_last_stats(out)
    keys = ["pairs", "WA", "WB", "WA_span", "WB_span", "truncated", "flagged", "row_passes", "panels", "units", "nchunks", "Kp", "ws_bytes", "peer_written"]
    d = {k: int(out[i]) for i, k in enumerate(keys)}
    d["moduli"] = d["pairs"]
    d["scheme"] = "residues"
    d["exact"] = d["pairs"] > 0 and d["truncated"] == 0     # windows cover the spans, no Inf/NaN: inner products exact, rounded once
    return d


def oz_last_mma_ms():
    """(summed ms, launches) of the tcgen05 kernel in the last tensor-path qgemm (CUDA events, blocks)."""
    n = C.c_int(0)
    ms = lib().qb_oz_last_mma_ms(C.byref(n))
    return float(ms), int(n.value)


def oz_last_mma_timeline(max_pairs=1024):
    """[(start_ms relative to the first launch, duration_ms)] of the tcgen05 kernel launches of the last tensor-path qgemm (blocks)."""
    buf = (C.c_double * (2 * max_pairs))()
    n = lib().qb_oz_last_mma_timeline(buf, max_pairs)
    return [(float(buf[2 * i]), float(buf[2 * i + 1])) for i in range(n)]


def oz_i8gemm(planesA, planesB, m, n, D, kb_begin=0, nkb=None):
    """The tcgen05 kernel alone: planes are torch int8 CUDA tensors [S][rows][Kp], D int32 [SA+SB-1][Mp][Np]."""
    SA, _, Kp = planesA.shape
    SB = planesB.shape[0]
    nkb = Kp // 128 - kb_begin if nkb is None else nkb
    check(lib().qb_oz_i8gemm_dev(C.c_void_p(planesA.data_ptr()), C.c_void_p(planesB.data_ptr()), SA, SB, m, n, Kp, kb_begin, nkb,
                                 C.c_void_p(D.data_ptr()), D.shape[1], D.shape[2], _stream()), "qb_oz_i8gemm_dev")
