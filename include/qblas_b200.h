/*
 * qblas_b200.h — C ABI of libqblas_b200.so: binary128 (IEEE quad) BLAS hot path on NVIDIA B200.
 *
 * Drop-in boundary for SwayamInSync/QBLAS ("QuadBLAS").  The reference is header-only: its C
 * entry points are `inline` definitions inside extern "C" in
 *   /root/reference/include/quadblas/interface/c_interface.hpp:15-146
 * and its C++ surface (QuadBLAS::gemm/gemv/dot/axpy, Vector<>, Matrix<>) calls the same free
 * functions.  This library exports
 *   (1) the reference C entry points under their exact names and signatures (section A), so an
 *       FFI consumer (ctypes / cffi / numpy-quaddtype style, c_interface.hpp:13) binds the same
 *       symbols, and
 *   (2) a quad-typed, device-pointer-aware, stream-aware extended API (section B) that the
 *       drop-in C++ headers in include/quadblas/ forward to and that bench.py / torch callers use.
 *
 * Data type: 16-byte little-endian IEEE-754 binary128 = SLEEF `Sleef_quad` = GCC `__float128`
 * (lo 64 mantissa bits first).  Only 16-byte element alignment is assumed (std::vector<Sleef_quad>,
 * /root/reference/test_quadblas.cpp:207).
 *
 * Pointers: every matrix/vector pointer may be a device pointer, managed memory, pinned host or
 * pageable host memory; it is classified with cudaPointerGetAttributes.  Host operands are staged
 * to the current CUDA device and results copied back before the call returns (calls are
 * synchronous, like the reference).  The *_dev entry points take device pointers only and are
 * asynchronous on the given stream.
 *
 * Numerical modes (qb_set_mode):
 *   QB_MODE_REFERENCE (default) reproduces the reference's reduction order bit for bit
 *       (SURVEY.md Appendix B): gemm k-panels of kc=126, dot chunked by quadblas_get_num_threads().
 *   QB_MODE_FAST is free to reorder; results satisfy |c^ - c| <= gamma_k (|A||B|)_ij,
 *       gamma_k = k u / (1 - k u), u = 2^-113.
 * There is no CPU fallback: every compute entry point fails with QB_ERR_CUDA when no sm_100 device
 * is usable.
 *
 * Environment (read when the library is loaded, so that an UNMODIFIED caller of the reference API can choose): QUADBLAS_MODE =
 * reference | fast, QUADBLAS_KC = k-panel of the reference-order qgemm (126), and OMP_NUM_THREADS = the default of
 * quadblas_get_num_threads() as in the reference (threading/openmp_utils.hpp:10-17: omp_get_max_threads()).
 *
 * Errors: the reference has no error channel (void/double returns, c_interface.hpp).  Here the
 * reference-named functions keep their signatures and record failures in a thread-local sticky
 * error (qb_last_error / qb_last_error_code); outputs are left untouched on failure, qdot/qnrm2
 * return NaN.  The qb_* functions return the code directly.  Nothing throws across the ABI.
 */
#ifndef QBLAS_B200_H
#define QBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb_quad { uint64_t lo, hi; } qb_quad; /* bit pattern of one binary128 */

enum { QB_OK = 0, QB_ERR_CUDA = 1, QB_ERR_ARG = 2, QB_ERR_ALLOC = 3 };
enum { QB_MODE_REFERENCE = 0, QB_MODE_FAST = 1 };

/* ------------------------------------------------------------------ A. reference C ABI */
/* c_interface.hpp:21  — returns (double)dot; full precision via qb_dot */
double quadblas_qdot(int n, void *x, int incx, void *y, int incy);
/* c_interface.hpp:34 */
double quadblas_qnrm2(int n, void *x, int incx);
/* c_interface.hpp:47 */
void quadblas_qaxpy(int n, double alpha, void *x, int incx, void *y, int incy);
/* c_interface.hpp:64  — trans in {T,t,C,c}: swap(m,n) and flip layout (lda unchanged) */
void quadblas_qgemv(char layout, char trans, int m, int n, double alpha, void *A, int lda, void *x, int incx,
                    double beta, void *y, int incy);
/* c_interface.hpp:95  — transa/transb accepted and IGNORED, exactly like the reference
 * (c_interface.hpp:109-112), unless qb_set_honor_trans(1) */
void quadblas_qgemm(char layout, char transa, char transb, int m, int n, int k, double alpha, void *A, int lda,
                    void *B, int ldb, double beta, void *C, int ldc);
/* c_interface.hpp:122,128 — T feeds the reference-order dot chunking (level1.hpp:46-65) */
void quadblas_set_num_threads(int num_threads);
int quadblas_get_num_threads(void);
/* c_interface.hpp:134 */
const char *quadblas_get_version(void);
/* c_interface.hpp:140 — 32-byte test (core/constants.hpp:18) */
int quadblas_is_aligned(const void *ptr);

/* ------------------------------------------------------------------ B. extended API */
int qb_init(void);                       /* create the context on the current CUDA device */
const char *qb_last_error(void);         /* thread-local sticky message ("" if none) */
int qb_last_error_code(void);
void qb_clear_error(void);
const char *qb_build_info(void);

void qb_set_mode(int mode);              /* QB_MODE_REFERENCE | QB_MODE_FAST */
int qb_get_mode(void);
void qb_set_kc(int kc);                  /* gemm k-panel (detail/blocking.hpp:21-66): 126 x86-64, 256 Apple */
int qb_get_kc(void);
void qb_set_honor_trans(int on);         /* extension: honour transa/transb in qgemm (default 0) */
int qb_get_honor_trans(void);

/* Fast-mode qgemm on the tensor cores (exact int8 residue planes + tcgen05 kind::i8 + Chinese-remainder reconstruction,
 * csrc/qb_ozaki.cu, csrc/qb_crt.cuh): 0 = never (integer-limb kernel only), 1 = automatic (m,n >= 128 and k >= 256; default),
 * 2 = whenever the planner accepts the operands.  Ignored in QB_MODE_REFERENCE.  Every finite or non-finite input is accepted
 * (the planner only declines when not even one pipeline unit fits the free device memory); then the integer-limb kernel runs. */
void qb_set_tensor_path(int v);
int qb_get_tensor_path(void);
/* Fast-mode accumulate of qdot / qnrm2 / qgemv: 2 (default) = large qgemv (both layouts) and large qnrm2 / qdot(x, x) on the FP64 pipe
 * (csrc/qslice.cuh: 22-bit slices as exact doubles, one rounding per result, rows it cannot guarantee recomputed by the window
 * kernel) and the window accumulator everywhere else; 1 = unrounded 192-bit window accumulator everywhere (csrc/qwide.cuh, one rounding per result);
 * 0 = chains of correctly rounded FMAs (the reference's per-element operation, level1.hpp:24, re-associated);
 * 3 = as 2 but the sliced kernels also take the sizes where they do not pay (any qgemv with n >= 128, any qnrm2): for tests.
 * Ignored in QB_MODE_REFERENCE. */
void qb_set_fast_variant(int v);
int qb_get_fast_variant(void);
/* Kernel of the reference-order qgemm (QuadBLAS::gemm, level3.hpp:215-336; same bits either way): 1 (default) = k_gemm_nb, staged
 * decoded operands and a branch-free step whose rare declined steps are redone out of line; 0 = k_gemm, the first version (side-by-side
 * in bench.py).  Also QBLAS_GEMM_KERNEL=0|1 in the environment when the library is loaded. */
void qb_set_ref_gemm_kernel(int v);
int qb_get_ref_gemm_kernel(void);
/* Pipelined all-host qgemm with beta = +-0 and contiguous rows of C: 1 (default) = C_in travels as one class byte per element
 * (finite >= +0 / finite with the sign bit / Inf-or-NaN) classified by host threads, and a kernel writes the stand-in +1 / -1 / NaN
 * on the device — beta * stand-in has exactly the bits of the reference's beta * C (level3.hpp:107), and 15/16 of the upload of C
 * is saved; 0 = upload C_in itself. */
void qb_set_beta0_classes(int v);
int qb_get_beta0_classes(void);
/* Rows of the last qb_gemv_dev that the sliced FP64 kernel declined and the window kernel recomputed (-1: that call did not take
 * the sliced path).  Synchronises the device: diagnostics and tests. */
int64_t qb_gemv_last_declined(void);
/* Row-pass hook of the device qgemm (qb_gemm_dev only).  When a callback is installed, the rows of C are produced in
 * min(min_passes, ceil(m / 128)) or more passes (tensor path; the integer-limb kernel makes one) and cb(row0, rows, user) runs on
 * the calling host thread right after the work that completes those rows has been enqueued on the stream: a collective issued
 * from the callback (e.g. an all-gather of those rows, ordered after the stream's work so far) overlaps the remaining work.
 * rows are relative to the C passed to the call ("m" direction: rows for row-major, also rows of op(A) for col-major).  Every row
 * is reported exactly once.  cb = NULL removes the hook.  Used by qblas_b200/dist.py. */
typedef void (*qb_pass_cb)(int64_t row0, int64_t rows, void *user);
void qb_set_gemm_pass_callback(qb_pass_cb cb, void *user, int min_passes);
/* Streamed B (qb_gemm_dev only, fast-mode tensor path): op(B) reaches the device in column panels of `panel_cols` columns (a multiple
 * of 256, or >= n) while the call is already computing — BASELINE config 4: the owner of B broadcasts it panel by panel and every rank
 * multiplies panel j while panel j+1 is still on the wire.  Before the library enqueues the first work that reads columns
 * [col0, col0 + cols) it calls cb(col0, cols, stream, &panel, &ld, user) on the calling host thread; the callee stores where that
 * panel lives (*panel: its column col0 is element 0 of each row for row-major B; *ld: leading dimension in elements; same layout
 * convention as the B argument) and makes `stream` (a cudaStream_t owned by the library) wait for the panel's arrival, e.g. with
 * cudaStreamWaitEvent.  d_colstats: device array of 3 n ints from qb_gemm_colstats_dev on the COMPLETE op(B) (the owner computes
 * it once and broadcasts these 12 n bytes first): the windows of all columns must be known before the first panel is reduced.
 * The B argument of the call is then not dereferenced.  cb = NULL removes the hook.  Non-zero return of cb fails the call. */
typedef int (*qb_bpanel_cb)(int64_t col0, int64_t cols, void *stream, const void **panel, int64_t *ld, void *user);
void qb_set_gemm_b_panels(qb_bpanel_cb cb, void *user, int64_t panel_cols, const void *d_colstats);
int qb_gemm_colstats_dev(char layout, char transb, int64_t k, int64_t n, const void *dB, int64_t ldb, void *d_colstats, void *stream);
/* B handed over as RESIDUE PLANES (row-sharded qgemm on N GPUs: every rank needs the residue planes of all of B, so the ranks compute
 * one column slice each and exchange the int8 planes over the NVSwitch multicast fabric instead of every rank reducing all of B;
 * qblas_b200/dist.py: qgemm_row_sharded(share_planes=...)).
 *   qb_crt_plan          the planner: spans of op(A) rows / op(B) columns (bits; from the statistics: max over non-empty lines of
 *                        emax + 113 - lmin) and k -> moduli count and windows.  Returns 0 when the windows cover the spans (exact),
 *                        a bit mask (1: A cut, 2: B cut) when they had to be capped, -1 when no plan exists.
 *   qb_crt_residues_dev  residue planes of the n columns of op(B) (k x n; same layout convention as qb_gemm_colstats_dev): plane i is
 *                        n x Kp int8 at d_planes + i * plane_stride (Kp = k rounded up to 128; plane_stride = 0: packed).  d_emax =
 *                        the first array of those columns' statistics.
 *   qb_set_gemm_b_planes the panels that the qb_set_gemm_b_panels callback hands over are such planes (`*panel` = plane 0 of the
 *                        panel's first column, `*ld` = bytes between planes), computed with `window` and at least the `moduli` the
 *                        call needs; moduli = 0 switches back to element panels.  The call fails (QB_ERR_CUDA, not supported) when
 *                        its own plan needs another window, more moduli, or the fix-up kernel (capped windows, Inf/NaN). */
int qb_crt_plan(int span_a, int span_b, int64_t k, int *moduli, int *window_a, int *window_b);
int qb_crt_residues_dev(char layout, char transb, int64_t k, int64_t n, const void *dB, int64_t ldb, const void *d_emax, int window, int moduli,
                        void *d_planes, int64_t plane_stride, void *stream);
void qb_set_gemm_b_planes(int moduli, int window);
/* Window of the tensor path.  Rows of op(A) / columns of op(B) are block fixed point with W_A / W_B-bit integers; the moduli must
 * cover W_A + W_B + log2(k) + 1 bits.  When the operands' bit spans fit into 2 x `bits` (every D53 / D113 input: 54 / 140 bits) the
 * windows are the spans and the inner products are exact.  Wider spans (exponent spreads of tens of binades inside one row) are
 * cut to the budget: low bits of the small elements of a row are dropped, every C element is tested against the dropped mass
 * (csrc/qb_crt.cuh: accept_msb) and the few that fail are recomputed in the window accumulator, so the result always satisfies the
 * fast-mode contract.  Default 144 (42 moduli at k = 8192), range 120..192. */
void qb_set_tensor_window(int bits);
int qb_get_tensor_window(void);
/* Pipeline unit of the tensor path: (rows of an A pass) x (columns of a B panel), default 2048 x 4096; rows / cols <= 0 restore the
 * default.  The tensor kernel of one unit overlaps the residues of the next pass / panel and the reconstruction of the previous unit. */
void qb_set_tensor_unit(int64_t rows, int64_t cols);
void qb_get_tensor_unit(int64_t *rows, int64_t *cols);
/* Rows of the FIRST pass and columns of the FIRST panel (0 = like the others): the residues of the first unit are the only ones
 * nothing can hide, so a short first unit lets the tensor kernel start early. */
void qb_set_tensor_ramp(int64_t rows, int64_t cols);
void qb_get_tensor_ramp(int64_t *rows, int64_t *cols);
/* Workspace of the tensor path (residue planes of the resident operand + double-buffered panels and residues): by default up to 85 % of
 * the free device memory, grow-only.  A limit > 0 caps it: the planner then keeps fewer passes / panels resident and sweeps the product
 * in blocks (the residues of the non-resident operand are recomputed per block), or shrinks the units; 0 restores the default. */
void qb_set_tensor_workspace_limit(size_t bytes);
size_t qb_get_tensor_workspace_limit(void);
/* Large all-host calls (quadblas_qgemm / quadblas_qgemv with host pointers) are pipelined: qgemm uploads the shared operand first,
 * then the rows stream in, are multiplied and stream out in `slabs` blocks on three streams; qgemv uploads A in `slabs` row blocks
 * while the earlier ones are multiplied.  Default 8 (1..16). */
void qb_set_host_slabs(int slabs);
int qb_get_host_slabs(void);
/* the row-pass partition helper: m rows in passes of at most `cap` rows (shape 1: short first and last pass); returns the number
 * of passes, out receives up to max_out sizes */
void qb_set_tensor_pass_shape(int shape);
int qb_get_tensor_pass_shape(void);
int qb_crt_pass_rows(int64_t m, int64_t cap, int shape, int64_t *out, int max_out);
/* plan of the last tensor-path qgemm (16 words): {moduli N (0: the tensor path did not run), W_A, W_B (window bits), widest span of
 * an A row, of a B column, truncation word (bit 0 / 1: the window of A / B is narrower than the span, bit 2: Inf / NaN present),
 * elements recomputed by the fix-up (waits for the call to finish), row passes, column panels, pipeline units, K chunks, padded K,
 * workspace bytes, peers written, 0, 0} */
void qb_oz_last_stats(int64_t *out16);
/* Summed device time (ms, CUDA events on the launching stream) of the tcgen05 kernel launches of
 * the last tensor-path qgemm; waits for them to finish.  *launches (optional) = how many. */
double qb_oz_last_mma_ms(int *launches);
/* profiling aid: (start relative to the first launch, duration) in ms of each of those launches, up to max_pairs; returns how many */
int qb_oz_last_mma_timeline(double *out, int max_pairs);
/* The tensor-core kernel alone (tests, the int8 peak microbenchmark): D[d] = sum_{s+t=d} A_s B_t^T over k-blocks
 * [kb_begin, kb_begin+nkb) of 128; planes are int8 [S][rows][Kp] (device), D is int32
 * [S_A+S_B-1][Mp][Np] with Mp % 128 == 0, Np % 256 == 0. */
int qb_oz_i8gemm_dev(const void *dPlanesA, const void *dPlanesB, int SA, int SB, int64_t m, int64_t n, int64_t Kp,
                     int64_t kb_begin, int64_t nkb, void *dD, int64_t Mp, int64_t Np, void *stream);

/* ---- peer memory: fused gather of a row-sharded qgemm (one process per GPU) ----
 * The reference has no multi-device code; BASELINE config 4 shards C by row blocks and gathers them.  Instead of a separate
 * all-gather that re-reads and re-sends each block, the reconstruction kernel of the residue-scheme tensor path can store every
 * finished element straight into the other GPUs' copies of C over NVLink:
 *   qb_peer_alloc / qb_peer_free      device memory that can be exported to the other processes of the node
 *   qb_peer_export / qb_peer_open     64-byte CUDA IPC handle of such a buffer / map a peer's buffer here (peer access is enabled)
 *   qb_set_gemm_peer_outputs(c, ptrs) for the following qb_gemm_dev calls: ptrs[q] is the address, inside peer q's mapped
 *                                     buffer, that corresponds to the C argument of the call (same strides); count 0 = off
 *   qb_get_gemm_peer_written()        how many peers the LAST qb_gemm_dev wrote (0 when it ran a path without the fused
 *                                     stores - the integer-limb kernel - so the caller must gather itself)
 * A peer address may also be an NVSwitch multicast address that maps the same buffer on every GPU of the node (one store then
 * reaches all of them: qblas_b200/dist.py SymmetricBuffer): pass count = 1.
 * Completion: the peers' copies are complete when the call's stream work has finished on EVERY rank (barrier after it). */
void *qb_peer_alloc(size_t bytes);
void qb_peer_free(void *p);
int qb_peer_export(void *p, void *handle64);
void *qb_peer_open(const void *handle64);
int qb_peer_close(void *p);
int qb_set_gemm_peer_outputs(int count, void *const *peer_C);
int qb_get_gemm_peer_written(void);

/* Quad-typed, host-or-device pointers, synchronous.  alpha/beta/result are HOST pointers to one
 * binary128 each.  Semantics = QuadBLAS::gemm/gemv/dot/axpy (level3.hpp:215, level2.hpp:85,
 * level1.hpp:80,190) and Vector::dot / Vector::norm (cpp_classes.hpp:66-81). */
int qb_gemm(char layout, char transa, char transb, int64_t m, int64_t n, int64_t k, const qb_quad *alpha,
            const void *A, int64_t lda, const void *B, int64_t ldb, const qb_quad *beta, void *C, int64_t ldc);
int qb_gemv(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *A, int64_t lda, const void *x,
            int64_t incx, const qb_quad *beta, void *y, int64_t incy);
int qb_dot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, qb_quad *result);
int qb_nrm2(int64_t n, const void *x, int64_t incx, qb_quad *result);
/* QuadBLAS::dot_kernel_vectorized (level1.hpp:14-35): the two-lane kernel over contiguous data in reference order, whatever
 * the mode and the thread count */
int qb_dot_kernel(int64_t n, const void *x, const void *y, qb_quad *result);
int qb_axpy(int64_t n, const qb_quad *alpha, const void *x, int64_t incx, void *y, int64_t incy);

/* Device pointers only, asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream).
 * d_result is a DEVICE pointer to 16 bytes.  No host synchronisation inside, with one exception: the fast-mode tensor path of
 * qb_gemm_dev waits once for its 3-integer plan (row / column bit spans) before it sizes the residue planes.  The tensor path runs on internal streams that are ordered
 * after `stream` at entry and that `stream` waits for before the call returns, so the caller sees plain stream order.
 * The library owns ONE grow-only workspace per process: calls issued on different streams must be ordered with respect to each
 * other by the caller (calls on one stream, or from one thread at a time on the default stream, always are). */
int qb_gemm_dev(char layout, char transa, char transb, int64_t m, int64_t n, int64_t k, const qb_quad *alpha,
                const void *dA, int64_t lda, const void *dB, int64_t ldb, const qb_quad *beta, void *dC,
                int64_t ldc, void *stream);
int qb_gemv_dev(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda,
                const void *dx, int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream);
/* The m rows are a block of a qgemv that has m_total rows in all (row blocks of a multi-GPU qgemv, slabs of a host matrix): the
 * fast-mode kernel choice and column splits follow m_total, so that every y_i carries the bits of the unsplit call. */
int qb_gemv_rows_dev(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda,
                     const void *dx, int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream, int64_t m_total);
int qb_dot_dev(int64_t n, const void *dx, int64_t incx, const void *dy, int64_t incy, void *d_result,
               void *stream);
int qb_nrm2_dev(int64_t n, const void *dx, int64_t incx, void *d_result, void *stream);
int qb_axpy_dev(int64_t n, const qb_quad *alpha, const void *dx, int64_t incx, void *dy, int64_t incy,
                void *stream);
/* Reference-order partials of a SHARD of a dot: the reference cuts [0,n) into T chunks of n/T
 * (level1.hpp:46-53); a rank that owns whole chunks calls this with its local vectors, the global
 * chunk length and its number of chunks (the last local chunk runs to n_local) and gets one
 * two-lane partial per chunk (unit strides) or one single-chain partial per chunk (strided). */
int qb_dot_partials_dev(int64_t n_local, const void *dx, int64_t incx, const void *dy, int64_t incy, int64_t chunk,
                        int64_t nchunks, void *d_partials, void *stream);
/* Combine `count` binary128 partials (device) in index order with add from +0 — the exchange step
 * of a sharded dot: partials are all-gathered as bytes, then folded on device. */
int qb_fold_partials_dev(int64_t count, const void *d_partials, int do_sqrt, void *d_result, void *stream);

/* Elementwise scalar ops on device arrays, for parity tests of the arithmetic core against
 * Sleef_{fma,mul,add,sqrt}q1_u05.  op: 0 fma (generic), 1 fma (chain form used by the kernels),
 * 2 mul, 3 add, 4 sqrt, 5 quad->double->quad round trip of casts. */
int qb_elementwise_dev(int op, int64_t n, const void *da, const void *db, const void *dc, void *dout,
                       void *stream);

/* Host buffers for QuadBLAS::aligned_alloc / aligned_free (memory/allocation.hpp:18-41): page-locked
 * (so staging copies run at full rate), at least 32-byte aligned; NULL on failure. */
void *qb_host_alloc(size_t bytes);
void qb_host_free(void *p);

/* Scalar casts (host, integer code; replace Sleef_cast_from_doubleq1 / Sleef_cast_to_doubleq1) */
qb_quad qb_from_double(double d);
double qb_to_double(qb_quad q);

/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t qb_launch_count(void);

/* register-resident qFMA throughput microbenchmark (the empirical integer-pipe roofline,
 * SURVEY.md §8d): runs `iters` dependent-chain steps on `ilp` independent accumulators per thread
 * over grid x block threads; returns total qFMAs issued through *n_fma.  d_sink: >=16 bytes. */
int qb_fma_microbench_dev(int variant, int blocks, int threads, int iters, void *d_sink, int64_t *n_fma,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QBLAS_B200_H */
