/* qb_internal.h — launcher prototypes shared by the kernel TUs and the C-ABI TU. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "q128.cuh"

namespace qb {

/* strides are in elements (16 B each) */
struct GemmArgs {
  int64_t m, n, k;
  q128 alpha, beta;
  const q128 *A; int64_t sai, sal;   /* op(A)(i,l) = A[i*sai + l*sal] */
  const q128 *B; int64_t sbl, sbj;   /* op(B)(l,j) = B[l*sbl + j*sbj] */
  q128 *C; int64_t sci, scj;         /* C(i,j)    = C[i*sci + j*scj] */
  int64_t kc;                        /* reference-order k-panel; >= k means a single chain */
#define QB_MAX_PEERS 8
  int npeer = 0;                     /* fused gather (residue-scheme tensor path only): C(i,j) is also stored to peerC[q][i*sci + j*scj] */
  q128 *peerC[QB_MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
cudaError_t launch_gemm(const GemmArgs &a, int mode, cudaStream_t st);

struct GemvArgs {
  int64_t m, n;                      /* y has m entries, x has n (after the C ABI's relabelling) */
  q128 alpha, beta;
  const q128 *A; int64_t lda;
  int col_major;                     /* 0: A[i*lda+j] (level2.hpp:15-50), 1: A[j*lda+i] (level2.hpp:53-82) */
  const q128 *x; int64_t incx;
  q128 *y; int64_t incy;
  q128 *work; int64_t work_elems;    /* device scratch (fast col-major split windows, the sliced kernel's tables and records), gemv_work_elems() */
  int64_t m_plan = 0;                /* > 0: these m rows are a slab of a qgemv with m_plan rows; the kernel choice follows m_plan (the column
                                        splits depend on n alone), so that a row gets the same bits whether the call is made in one piece or in slabs */
};
int64_t gemv_work_elems(int64_t m, int64_t n, int col_major, int mode, int64_t m_plan = 0);
const uint8_t *gemv_sliced_rowflags(const q128 *work, int64_t m, int64_t n, int col_major, int64_t m_plan = 0);   /* the declined-row flags of the sliced row-major kernel inside its scratch */
cudaError_t launch_gemv(const GemvArgs &a, int mode, cudaStream_t st);

struct DotArgs {
  int64_t n;
  const q128 *x; int64_t incx;
  const q128 *y; int64_t incy;
  int T;                             /* reference-order chunk count (quadblas_get_num_threads) */
  int do_sqrt;                       /* nrm2 */
  q128 *result;                      /* device, 16 B */
  q128 *work; int64_t work_elems;    /* device scratch for partials */
  unsigned *ticket = nullptr;        /* device, zero between calls: the last-CTA-done counter of the fast one-launch reduction */
  unsigned *only_if = nullptr;       /* device word: the sliced qnrm2 kernel sets it when it declines; the window kernel queued behind it runs only then */
};
int64_t dot_work_elems(int64_t n, int T, int mode);
cudaError_t launch_dot(const DotArgs &a, int mode, cudaStream_t st);
cudaError_t launch_dot_partials(const DotArgs &a, int64_t chunk, int nchunks, q128 *partials, cudaStream_t st);
cudaError_t launch_fold(int64_t count, const q128 *partials, int do_sqrt, q128 *result, cudaStream_t st);

cudaError_t launch_axpy(int64_t n, q128 alpha, const q128 *x, int64_t incx, q128 *y, int64_t incy, cudaStream_t st);
cudaError_t launch_elementwise(int op, int64_t n, const q128 *a, const q128 *b, const q128 *c, q128 *out, cudaStream_t st);
cudaError_t launch_fma_microbench(int variant, int blocks, int threads, int iters, q128 *sink, int64_t *n_fma, cudaStream_t st);

void count_launch(int n = 1);
int ref_gemm_kernel(); /* qb_set_ref_gemm_kernel: 1 = k_gemm_nb (branch-free step, default), 0 = k_gemm (first version) */
int fast_variant();   /* qb_set_fast_variant: 1 = window accumulator (default), 0 = rounded-FMA chains */

/* ---- fast-mode tensor-core GEMM (qb_ozaki.cu) ---- */
#define QB_OZ_MAX_SLICES 24
#define QB_MAX_DEVICES 64
struct OzStats {
  int SA = 0, SB = 0, ndiag = 0, nchunks = 0, row_passes = 0; int64_t pairs = 0, ws_bytes = 0, Kp = 0; int keep = 0; int64_t flagged = 0; int redo_passes = 0;
  int scheme = 0, WA = 0, WB = 0; int peer_written = 0;
  int panels = 0; int64_t units = 0; int WA_nat = 0, WB_nat = 0; int truncated = 0;   /* bit 0 / 1: the window of A / B is narrower than the span; bit 2: Inf / NaN present */
};
/* row-pass hook: when set, the C rows are produced in at least `min_passes` passes and cb(row0, rows, user) is called on the
 * host after the work that completes those rows has been ENQUEUED on the stream (so a collective issued from the callback
 * overlaps the remaining work) */
typedef void (*oz_pass_cb)(int64_t row0, int64_t rows, void *user);
/* streamed B: called on the host before the library enqueues the first work that reads columns [col0, col0 + cols) of op(B);
 * the callee returns where that panel lives (*panel, leading dimension *ld in elements, same layout as the B argument) and makes
 * `stream` (a cudaStream_t of the library) wait for the panel's arrival.  non-zero return = failure. */
typedef int (*oz_bpanel_cb)(int64_t col0, int64_t cols, void *stream, const void **panel, int64_t *ld, void *user);
struct OzHooks {
  oz_pass_cb cb = nullptr; void *cb_user = nullptr; int min_passes = 1;
  oz_bpanel_cb bp = nullptr; void *bp_user = nullptr; int64_t bp_cols = 0;
  const int *bstats = nullptr;       /* device: column statistics of op(B) from launch_colstats (3 n ints); skips the scan of B */
  /* the panels bp hands over are RESIDUE PLANES of op(B) (N planes of cols x Kp int8, `ld` returned by bp = bytes between planes),
   * computed with window bplanes_W and bplanes_N moduli: no residues of B are computed here */
  int bplanes_N = 0, bplanes_W = 0;
  /* streamed rows (the all-host path): called once per row pass before its first use; the callee makes sA wait for the rows of A
   * and sF for the rows of C_in.  non-zero return = failure. */
  int (*rows_in)(int64_t row0, int64_t rows, void *sA, void *sF, void *user) = nullptr; void *rows_user = nullptr;
  int order = 0;                     /* 0: panels outer (A planes resident), 1: passes outer (B planes resident, rows complete pass by pass) */
};
/* *used = 0: the planner declined (no TMA entry point, no workspace) and nothing was written */
cudaError_t launch_gemm_ozaki(const GemmArgs &a, cudaStream_t st, int *used, const OzHooks &h);
void launch_crt_residues(const q128 *X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *emax, int W, int N, int64_t Kp, int8_t *planes,
                         cudaStream_t st, int64_t pstride);
cudaError_t oz_prepare_device();
cudaError_t launch_colstats(const q128 *B, int64_t n, int64_t k, int64_t sbj, int64_t sbl, int *stats, cudaStream_t st);
cudaError_t launch_oz_mma(const int8_t *pA, const int8_t *pB, int SA, int SB, int64_t m, int64_t n, int64_t Kp, int kb_begin, int nkb,
                          int32_t *D, int64_t Mp, int64_t Np, cudaStream_t st, int keep = 0);
OzStats oz_last_stats(bool wait = true);   /* wait: block until the fix-up count of the last call has arrived */
std::vector<int64_t> oz_crt_pass_rows(int64_t m, int64_t cap, int shape);
void oz_set_pass_shape(int v);
int oz_get_pass_shape();
void oz_set_window(int bits);  /* bits per operand window the planner grants when the spans do not fit the moduli (default 144) */
int oz_get_window();
void oz_set_unit(int64_t rows, int64_t cols);   /* pipeline unit: rows of an A pass x columns of a B panel (default 2048 x 4096) */
void oz_get_unit(int64_t *rows, int64_t *cols);
void oz_set_ramp(int64_t rows, int64_t cols);    /* rows of the first pass / columns of the first panel (0 = like the others) */
void oz_get_ramp(int64_t *rows, int64_t *cols);
void oz_set_ws_limit(size_t bytes);              /* cap of the tensor path's workspace (0 = 85 % of the free device memory) */
size_t oz_get_ws_limit();
double oz_last_mma_ms(int *launches);
int oz_last_mma_timeline(double *out, int max_pairs);
void oz_release();

} // namespace qb
