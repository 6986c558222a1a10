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

/* Fast-mode qgemm on the tensor cores (exact int8 slicing + tcgen05 kind::i8, csrc/qb_ozaki.cu):
 * 0 = never (integer-limb kernel only), 1 = automatic (m,n >= 128 and k >= 256; default),
 * 2 = whenever the planner accepts the operands.  Ignored in QB_MODE_REFERENCE. */
void qb_set_tensor_path(int v);
int qb_get_tensor_path(void);
/* Fast-mode accumulate of qdot / qnrm2 / qgemv: 1 (default) = unrounded 192-bit window accumulator
 * (csrc/qwide.cuh, one rounding per result), 0 = chains of correctly rounded FMAs (the
 * reference's per-element operation, level1.hpp:24, re-associated).  Ignored in QB_MODE_REFERENCE. */
void qb_set_fast_variant(int v);
int qb_get_fast_variant(void);
/* Row-pass hook of the device qgemm (qb_gemm_dev).  When a callback is installed, the rows of C are produced in at
 * least `min_passes` passes (tensor path; the integer-limb kernel makes one) and cb(row0, rows, user) runs on the
 * calling host thread right after the work of each pass has been enqueued on the stream: a collective issued from the
 * callback (e.g. an all-gather of those rows, ordered after the stream's work so far) overlaps the next pass.  rows are
 * relative to the C passed to the call ("m" direction: rows for row-major, also rows of op(A) for col-major).  Every row
 * is reported exactly once.  cb = NULL removes the hook.  Used by qblas_b200/dist.py. */
typedef void (*qb_pass_cb)(int64_t row0, int64_t rows, void *user);
void qb_set_gemm_pass_callback(qb_pass_cb cb, void *user, int min_passes);
/* Accuracy setting of the tensor path.  keep = 0: every digit-plane product is computed, the inner
 * products are EXACT and rounded once.  keep = d > 0 (default 16): only the d most significant
 * diagonals are multiplied; every element is checked (|J| >= 2^125, csrc/qb_ozaki.cu) and the few that
 * fail are recomputed in the window accumulator, so the result always satisfies the fast-mode
 * contract |c^ - c| <= gamma_k (|A||B|)_ij, but it is no longer the exact sum rounded once. */
void qb_set_tensor_keep(int keep);
int qb_get_tensor_keep(void);
/* How the tensor path forms the exact integer inner products.  1 (default) = residue scheme
 * (csrc/qb_crt.cuh): the block-fixed-point integers are reduced modulo N pairwise coprime moduli <= 256,
 * ONE int8 GEMM per modulus, exact Chinese-remainder reconstruction, one rounding - always exact, so
 * qb_set_tensor_keep does not apply; N = 41 for full 113-bit mantissas at k = 8192 (vs 324 / 136 digit-plane
 * products).  0 = digit diagonals (csrc/qb_ozaki.cu).  The residue scheme hands over to the digit
 * diagonals when its moduli cannot cover the operands' bit span (W_A + W_B + log2 k + 1 > 341) or k > 65536. */
void qb_set_tensor_scheme(int scheme);
int qb_get_tensor_scheme(void);
/* Row passes of the residue scheme (they are software-pipelined: tensor kernel of pass p || residues of pass p+1 || fold of
 * pass p-1).  0 (default, the measured setting): equal passes.  1 (experimental, not yet measured on hardware): a short first
 * and a short last pass, whose residues / fold cannot hide behind a tensor pass.  Ignored while a row-pass callback is set.
 * qb_crt_pass_rows reports the partition the library would use for m rows with at most `cap` rows per pass (returns the
 * number of passes; out receives up to max_out sizes). */
/* Large all-host qgemm calls (quadblas_qgemm / qb_gemm with host pointers) are pipelined: the shared operand is uploaded first,
 * then C is cut into `slabs` blocks whose uploads, compute and downloads overlap on three streams.  Default 4 (the measured
 * setting); more slabs shorten the tail after the last upload (1..16). */
void qb_set_host_slabs(int slabs);
int qb_get_host_slabs(void);
void qb_set_tensor_pass_shape(int shape);
int qb_get_tensor_pass_shape(void);
int qb_crt_pass_rows(int64_t m, int64_t cap, int shape, int64_t *out, int max_out);
/* plan of the last tensor-path qgemm: {S_A, S_B, diagonals, K chunks, row passes, digit-plane products
 * per row pass, workspace bytes, padded K, diagonals kept, elements sent to the fix-up, row passes
 * redone with all diagonals, residue-scheme word}.  Residue scheme: S_A / S_B = bytes of the widest row /
 * column integer, "diagonals" = "products" = N moduli, last word = 1 | W_A << 8 | W_B << 24 (bit spans);
 * digit diagonals: last word = 0. */
void qb_oz_last_stats(int64_t *out12);
/* Summed device time (ms, CUDA events on the launching stream) of the tcgen05 kernel launches of
 * the last tensor-path qgemm; waits for them to finish.  *launches (optional) = how many. */
double qb_oz_last_mma_ms(int *launches);
/* The tensor-core kernel alone (tests/profiling): D[d] = sum_{s+t=d} A_s B_t^T over k-blocks
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
 *                                     stores - digit diagonals, integer-limb kernel - so the caller must gather itself)
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
int qb_axpy(int64_t n, const qb_quad *alpha, const void *x, int64_t incx, void *y, int64_t incy);

/* Device pointers only, asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream).
 * d_result is a DEVICE pointer to 16 bytes.  No host synchronisation inside, with one exception: the fast-mode tensor path of
 * qb_gemm_dev waits once for its 3-integer plan (row / column bit spans) before it sizes the residue planes; the bounded
 * digit-diagonal setting also reads one counter per row pass.  The tensor path runs on internal streams that are ordered
 * after `stream` at entry and that `stream` waits for before the call returns, so the caller sees plain stream order.
 * The library owns ONE grow-only workspace per process: calls issued on different streams must be ordered with respect to each
 * other by the caller (calls on one stream, or from one thread at a time on the default stream, always are). */
int qb_gemm_dev(char layout, char transa, char transb, int64_t m, int64_t n, int64_t k, const qb_quad *alpha,
                const void *dA, int64_t lda, const void *dB, int64_t ldb, const qb_quad *beta, void *dC,
                int64_t ldc, void *stream);
int qb_gemv_dev(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda,
                const void *dx, int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream);
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
