/*
 * qb_abi.cu — the C ABI of libqblas_b200.so (see include/qblas_b200.h).
 *
 * Host-side mirror of the reference's public surface:
 *   quadblas_q{dot,nrm2,axpy,gemv,gemm}, set/get_num_threads, get_version, is_aligned
 *     (/root/reference/include/quadblas/interface/c_interface.hpp:21-146) — same names, argument
 *     meaning and marshalling (double alpha/beta widened exactly, layout char, gemv transpose by
 *     relabelling :79-85, gemm transa/transb ignored :109-112);
 *   qb_* = QuadBLAS::gemm/gemv/dot/axpy with quad-typed scalars (level3.hpp:215, level2.hpp:85,
 *     level1.hpp:80,190) for the drop-in C++ headers and for device-resident callers.
 * There is no CPU compute path here: if CUDA is unusable the calls fail (QB_ERR_CUDA).
 */
#include "../../include/qblas_b200.h"
#include "qb_internal.h"
#include "qb_crt.cuh"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <cstdlib>
#include <strings.h>
#include <limits>

namespace qb {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_mode{QB_MODE_REFERENCE};
static std::atomic<int> g_kc{126};
static std::atomic<int> g_honor_trans{0};
/* hooks of the device qgemm (qb_set_gemm_pass_callback / qb_set_gemm_b_panels / qb_set_gemm_peer_outputs); used under the library mutex */
static qb_pass_cb g_pass_cb = nullptr;
static void *g_pass_user = nullptr;
static int g_pass_min = 1;
static qb_bpanel_cb g_bp_cb = nullptr;
static void *g_bp_user = nullptr;
static int64_t g_bp_cols = 0;
static const int *g_bp_stats = nullptr;
static int g_bp_planes_N = 0, g_bp_planes_W = 0;   /* the panels are residue planes (qb_set_gemm_b_planes) */
static int g_npeer = 0;
static void *g_peer[QB_MAX_PEERS];
static int g_peer_written = 0;
static std::atomic<int> g_host_slabs{8}; /* C slabs of the pipelined all-host qgemm / row slabs of the all-host qgemv */
static std::atomic<int> g_tensor{1};  /* fast-mode tensor path: 0 off, 1 auto (size threshold), 2 always */
static std::atomic<int> g_beta0_classes{1};   /* qb_set_beta0_classes: the pipelined host qgemm sends C_in as class bytes when beta = 0 */
static std::atomic<int> g_fastvar{2}; /* fast-mode level-1/2 accumulate: 2 = sliced FP64 accumulate where it applies (large row-major qgemv), window
                                         accumulator elsewhere; 1 = window accumulator everywhere; 0 = rounded-FMA chains; 3 = 2 without the size thresholds (tests) */
int fast_variant() { return g_fastvar.load(); }
static std::atomic<int> g_gemm_kernel{1}; /* reference-order qgemm: 1 = k_gemm_nb, 0 = k_gemm (qb_set_ref_gemm_kernel / QBLAS_GEMM_KERNEL) */
int ref_gemm_kernel() { return g_gemm_kernel.load(); }
static std::atomic<int> g_threads{0}; /* 0 = not set -> OMP_NUM_THREADS, then hardware concurrency (omp_get_max_threads) */

/* Environment, read once when the library is loaded, so that an UNMODIFIED caller of the reference API can choose the
 * numerical mode without a source change:
 *   QUADBLAS_MODE = reference | fast      (default reference: bit for bit the reference's results)
 *   QUADBLAS_KC   = k-panel of the reference-order qgemm (default 126, detail/blocking.hpp:46; 256 mimics Apple Silicon)
 *   OMP_NUM_THREADS is read where the reference reads it: the default of quadblas_get_num_threads()
 *   (threading/openmp_utils.hpp:10-17: omp_get_max_threads()), which defines the chunking of the reference-order dot. */
static int env_threads()
{
  const char *v = getenv("OMP_NUM_THREADS");
  if (!v || !*v) return 0;
  const long t = strtol(v, nullptr, 10);     /* "a,b,c" (nested levels): the first entry is the outermost team size */
  return t > 0 && t < (1 << 20) ? (int)t : 0;
}
struct EnvInit {
  EnvInit()
  {
    if (const char *v = getenv("QUADBLAS_MODE")) {
      if (!strcasecmp(v, "fast") || !strcmp(v, "1")) g_mode.store(QB_MODE_FAST);
      else if (!strcasecmp(v, "reference") || !strcasecmp(v, "ref") || !strcmp(v, "0")) g_mode.store(QB_MODE_REFERENCE);
    }
    if (const char *v = getenv("QBLAS_GEMM_KERNEL")) g_gemm_kernel.store(v[0] == '0' ? 0 : 1);
    if (const char *v = getenv("QUADBLAS_KC")) { const long kc = strtol(v, nullptr, 10); if (kc > 0 && kc < (1 << 30)) g_kc.store((int)kc); }
  }
};
static EnvInit g_env_init;

static thread_local int t_err_code = 0;
static thread_local char t_err_msg[512] = "";

static int fail(int code, const char *what, cudaError_t ce = cudaSuccess)
{
  t_err_code = code;
  if (ce != cudaSuccess) snprintf(t_err_msg, sizeof t_err_msg, "%s: %s", what, cudaGetErrorString(ce));
  else snprintf(t_err_msg, sizeof t_err_msg, "%s", what);
  return code;
}
static void clear_error() { t_err_code = 0; t_err_msg[0] = 0; }

static int num_threads()
{
  int t = g_threads.load();
  if (t <= 0) t = env_threads();
  if (t <= 0) { t = (int)std::thread::hardware_concurrency(); if (t <= 0) t = 1; }
  return t;
}

/* ---- per-device scratch (dot partials, staged scalars, staging buffers, streams of the pipelined host paths); one mutex
 * guards all of it.  ensure_device() selects the entry of the current CUDA device: a process that switches devices between
 * calls gets an independent set per device. ---- */
struct Scratch {
  bool ready = false;
  q128 *work = nullptr; int64_t work_elems = 0;
  q128 *result = nullptr;
  void *stage[3] = {nullptr, nullptr, nullptr}; size_t stage_bytes[3] = {0, 0, 0};
  cudaStream_t cs = nullptr, ks = nullptr, ds = nullptr;   /* copy-in, kernels, copy-out */
  uint8_t *codes_h = nullptr, *codes_d = nullptr; size_t codes_bytes = 0;   /* one class byte per element of C_in (beta = 0 host path) */
};
static std::recursive_mutex g_mu;
static Scratch g_scr[QB_MAX_DEVICES];
static int g_curdev = 0;
static inline Scratch &S() { return g_scr[g_curdev]; }
static struct { const uint8_t *flags = nullptr; int64_t m = 0; int dev = 0; } g_last_sliced;   /* qb_gemv_last_declined */

static int ensure_device()
{
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "no usable CUDA device (cudaGetDevice)", e);
  if (dev < 0 || dev >= QB_MAX_DEVICES) return fail(QB_ERR_CUDA, "device ordinal out of range");
  g_curdev = dev;
  Scratch &s = g_scr[dev];
  if (!s.ready) {
    e = cudaMalloc((void **)&s.result, 64);          /* 16 B staged result | 16 B spare | the dot ticket (zero between calls) */
    if (e == cudaSuccess) e = cudaMemset(s.result, 0, 64);
    if (e != cudaSuccess) return fail(QB_ERR_CUDA, "cudaMalloc(result)", e);
    s.ready = true;
  }
  return QB_OK;
}

/* the three streams of the pipelined host paths, created together (all or none) */
static int ensure_streams()
{
  Scratch &s = S();
  if (s.cs) return QB_OK;
  cudaStream_t t[3] = {nullptr, nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&t[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    for (int i = 0; i < 3; ++i) if (t[i]) cudaStreamDestroy(t[i]);
    return fail(QB_ERR_CUDA, "stream creation", e);
  }
  s.cs = t[0]; s.ks = t[1]; s.ds = t[2];
  return QB_OK;
}

static int ensure_work(int64_t elems)
{
  Scratch &s = S();
  if (elems <= s.work_elems) return QB_OK;
  cudaFree(s.work);
  s.work = nullptr; s.work_elems = 0;
  cudaError_t e = cudaMalloc((void **)&s.work, (size_t)elems * 16);
  if (e != cudaSuccess) return fail(QB_ERR_ALLOC, "cudaMalloc(work)", e);
  s.work_elems = elems;
  return QB_OK;
}

static bool is_device_ptr(const void *p)
{
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

/* ---- pageable host memory <-> device at DMA speed ----
 * cudaMemcpy from / to pageable memory is staged by the driver through a small pinned buffer by ONE thread (~10 GB/s, a fifth of
 * the PCIe rate), and that is what std::vector<Sleef_quad> / numpy callers of the reference API hand over (the README's 1000^3
 * call moves 64 MB around 0.7 ms of compute).  Here a few host threads copy 4 MB chunks between the caller's memory and a ring
 * of page-locked slots while the copy engine moves the previous chunks, so the transfer runs at the memcpy rate of several
 * cores.  Only for unregistered host memory and transfers of 8 MB and more; everything else is a plain cudaMemcpy. */
static constexpr size_t PG_CHUNK = (size_t)4 << 20;
static constexpr int PG_THREADS = 4, PG_SLOTS = 2;
struct PageRing {
  char *base = nullptr; bool tried = false;
  int dev = -1;                                   /* the streams / events below belong to this device */
  cudaStream_t st[PG_THREADS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[PG_THREADS][PG_SLOTS] = {};
};
static PageRing g_rings[2];   /* guarded by g_mu like every host path; [0] uploads (and every synchronous path), [1] the download thread of the pipelined qgemm */
static bool is_pageable(const void *p)
{
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}
static bool ring_ready(int dev, int ring)
{
  PageRing &g_ring = g_rings[ring];
  if (!g_ring.base) {
    if (g_ring.tried) return false;
    g_ring.tried = true;
    void *p = nullptr;
    if (cudaHostAlloc(&p, PG_CHUNK * PG_THREADS * PG_SLOTS, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    g_ring.base = (char *)p;
  }
  if (g_ring.dev != dev) {   /* (re)create the per-thread streams and per-slot events on the current device */
    for (int w = 0; w < PG_THREADS; ++w) {
      if (g_ring.st[w]) { cudaStreamDestroy(g_ring.st[w]); g_ring.st[w] = nullptr; }
      for (int i = 0; i < PG_SLOTS; ++i) if (g_ring.ev[w][i]) { cudaEventDestroy(g_ring.ev[w][i]); g_ring.ev[w][i] = nullptr; }
    }
    g_ring.dev = -1;
    for (int w = 0; w < PG_THREADS; ++w) {
      if (cudaStreamCreateWithFlags(&g_ring.st[w], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
      for (int i = 0; i < PG_SLOTS; ++i)
        if (cudaEventCreateWithFlags(&g_ring.ev[w][i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    g_ring.dev = dev;
  }
  return true;
}
/* to_dev: host -> device, else device -> host.  Synchronous. */
static cudaError_t paged_copy(void *dst, const void *src, size_t bytes, bool to_dev, int ring = 0)
{
  PageRing &g_ring = g_rings[ring];
  const void *host = to_dev ? src : dst;
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes < 2 * PG_CHUNK || !is_pageable(host) || !ring_ready(dev, ring))
    return cudaMemcpy(dst, src, bytes, to_dev ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
  const size_t nchunks = (bytes + PG_CHUNK - 1) / PG_CHUNK;
  cudaError_t errs[PG_THREADS];
  std::thread th[PG_THREADS];
  int started = 0;
  auto worker = [&](int w) {
      cudaError_t e = cudaSetDevice(dev);
      const cudaStream_t st = g_ring.st[w];
      cudaEvent_t *ev = g_ring.ev[w];
      size_t pending_off[PG_SLOTS] = {0, 0}, pending_len[PG_SLOTS] = {0, 0};
      int it = 0;
      for (size_t c = (size_t)w; c < nchunks && e == cudaSuccess; c += PG_THREADS, ++it) {
        const int slot = it % PG_SLOTS;
        char *pin = g_ring.base + ((size_t)w * PG_SLOTS + slot) * PG_CHUNK;
        const size_t off = c * PG_CHUNK, len = std::min(PG_CHUNK, bytes - off);
        if (it >= PG_SLOTS) {      /* the slot's previous transfer must be over */
          e = cudaEventSynchronize(ev[slot]);
          if (e == cudaSuccess && !to_dev) memcpy((char *)dst + pending_off[slot], pin, pending_len[slot]);
        }
        if (e != cudaSuccess) break;
        if (to_dev) {
          memcpy(pin, (const char *)src + off, len);
          e = cudaMemcpyAsync((char *)dst + off, pin, len, cudaMemcpyHostToDevice, st);
        } else {
          e = cudaMemcpyAsync(pin, (const char *)src + off, len, cudaMemcpyDeviceToHost, st);
          pending_off[slot] = off; pending_len[slot] = len;
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev[slot], st);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess && !to_dev) {   /* drain the slots still holding downloaded chunks */
        const int done = it;
        for (int b = std::max(0, done - PG_SLOTS); b < done; ++b) {
          const int slot = b % PG_SLOTS;
          memcpy((char *)dst + pending_off[slot], g_ring.base + ((size_t)w * PG_SLOTS + slot) * PG_CHUNK, pending_len[slot]);
        }
      }
      errs[w] = e;
  };
  for (int w = 0; w < PG_THREADS; ++w) errs[w] = cudaSuccess;
  for (int w = 1; w < PG_THREADS; ++w) {
    try { th[w] = std::thread(worker, w); ++started; }
    catch (...) { break; }             /* no thread to be had: nothing may throw across the C ABI; the caller's thread does that share too */
  }
  worker(0);
  for (int w = started + 1; w < PG_THREADS; ++w) worker(w);
  cudaError_t e = cudaSuccess;
  for (int w = 1; w <= started; ++w) th[w].join();
  for (int w = 0; w < PG_THREADS; ++w) if (errs[w] != cudaSuccess) e = errs[w];
  return e;
}

/* staging slot: returns a device pointer holding `bytes` copied from host `src` (or src itself if
 * it is already device-accessible) */
static int stage_in(int slot, const void *src, size_t bytes, bool copy, const void **dptr, bool *staged)
{
  if (bytes == 0 || is_device_ptr(src)) { *dptr = src; *staged = false; return QB_OK; }
  Scratch &s = S();
  if (s.stage_bytes[slot] < bytes) {
    cudaFree(s.stage[slot]); s.stage[slot] = nullptr; s.stage_bytes[slot] = 0;
    cudaError_t e = cudaMalloc(&s.stage[slot], bytes);
    if (e != cudaSuccess) return fail(QB_ERR_ALLOC, "cudaMalloc(staging)", e);
    s.stage_bytes[slot] = bytes;
  }
  if (copy) {
    cudaError_t e = paged_copy(s.stage[slot], src, bytes, true);
    if (e != cudaSuccess) return fail(QB_ERR_CUDA, "cudaMemcpy H2D", e);
  }
  *dptr = s.stage[slot];
  *staged = true;
  return QB_OK;
}

static inline q128 toq(const qb_quad *p) { q128 r; r.lo = p->lo; r.hi = p->hi; return r; }
static inline bool is_col(char layout) { return layout == 'C' || layout == 'c'; }
static inline bool is_trans(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }

/* extent in elements of a strided vector / matrix footprint */
static inline size_t vec_bytes(int64_t n, int64_t inc) { return n <= 0 ? 0 : (size_t)((n - 1) * inc + 1) * 16; }
static inline size_t mat_bytes(int64_t outer, int64_t inner, int64_t ld) { return (outer <= 0 || inner <= 0) ? 0 : (size_t)((outer - 1) * ld + inner) * 16; }

/* ---- beta = 0 on the pipelined host path: C_in by its class, not by its bytes ----------------------------------------------------
 * The reference evaluates beta * C even for beta = 0 (level3.hpp:107), so C_in cannot simply be ignored: 0 * C is +-0 with the sign
 * of beta XOR the sign of C for a finite C, and NaN for an Inf or NaN.  But nothing else of C_in reaches the result.  So instead of
 * uploading 16 bytes per element, host threads classify C_in (one byte per element: 0 = finite >= +0, 1 = finite with the sign bit,
 * 2 = Inf / NaN) while the shared operand is on the wire, the class bytes are uploaded, and a kernel writes the stand-in +1, -1 or
 * NaN into the device copy of C: beta * stand-in has exactly the bits of beta * C_in, whatever the mode.  An 8192 x 8192 C: 64 MB
 * instead of 1 GiB over PCIe.  Used when the rows of C are contiguous (ldc = row length: the download then restores no padding). */
struct CodeScan {
  const q128 *C = nullptr;
  uint8_t *codes = nullptr;
  int64_t count = 0, chunk = 1 << 16, nchunks = 0;
  std::atomic<int64_t> next{0};
  std::vector<std::atomic<uint8_t>> done;
  std::vector<std::thread> th;
  explicit CodeScan(int64_t nch) : done((size_t)nch) { for (auto &d : done) d.store(0, std::memory_order_relaxed); }
  void work()
  {
    for (;;) {
      const int64_t c = next.fetch_add(1, std::memory_order_relaxed);
      if (c >= nchunks) return;
      const int64_t lo = c * chunk, hi = std::min(count, lo + chunk);
      for (int64_t i = lo; i < hi; ++i) {
        const uint64_t h = C[i].hi;
        codes[i] = ((h >> 48) & 0x7fffu) == 0x7fffu ? (uint8_t)2 : (uint8_t)(h >> 63);
      }
      done[(size_t)c].store(1, std::memory_order_release);
    }
  }
  void wait_elems(int64_t lo, int64_t hi)   /* until the classes of [lo, hi) are written */
  {
    for (int64_t c = lo / chunk; c <= (hi - 1) / chunk && c < nchunks; ++c)
      while (!done[(size_t)c].load(std::memory_order_acquire)) std::this_thread::yield();
  }
  ~CodeScan() { for (auto &t : th) if (t.joinable()) t.join(); }
};

__global__ void k_c_standin(const uint8_t *codes, q128 *C, int64_t count)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t c = codes[i];
    q128 v;
    v.lo = 0;
    v.hi = c == 2 ? 0x7fff800000000000ull : (0x3fff000000000000ull | ((uint64_t)(c & 1) << 63));
    C[i] = v;
  }
}

/* a handful of ordering events for one pipelined host call; destroyed on scope exit */
struct EventSet {
  std::vector<cudaEvent_t> ev;
  bool make(int count)
  {
    while ((int)ev.size() < count) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
      ev.push_back(e);
    }
    return true;
  }
  ~EventSet() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};

/* (layout, transposes, leading dimensions) -> the stride form the kernels take */
static GemmArgs make_gemm_args(char layout, char ta, char tb, int64_t m, int64_t n, int64_t k, q128 alpha, const void *dA, int64_t lda, const void *dB,
                               int64_t ldb, q128 beta, void *dC, int64_t ldc)
{
  const bool col = is_col(layout);
  const bool honor = g_honor_trans.load() != 0;
  const bool tA = honor && is_trans(ta), tB = honor && is_trans(tb);
  GemmArgs g;
  g.m = m; g.n = n; g.k = k; g.alpha = alpha; g.beta = beta;
  g.A = (const q128 *)dA; g.B = (const q128 *)dB; g.C = (q128 *)dC;
  /* element (i,l) of op(A): row-walk storage iff (col != tA) is false */
  if (col != tA) { g.sai = 1; g.sal = lda; } else { g.sai = lda; g.sal = 1; }
  if (col != tB) { g.sbl = 1; g.sbj = ldb; } else { g.sbl = ldb; g.sbj = 1; }
  if (col) { g.sci = 1; g.scj = ldc; } else { g.sci = ldc; g.scj = 1; }
  g.kc = g_kc.load();
  return g;
}
/* C^T = op(B)^T op(A)^T: the same product with the roles of the operands exchanged (rows of the new problem = columns of C) */
static void swap_roles(GemmArgs &g)
{
  std::swap(g.m, g.n);
  const q128 *A0 = g.A; const int64_t sai0 = g.sai, sal0 = g.sal;
  g.A = g.B; g.sai = g.sbj; g.sal = g.sbl;
  g.B = A0; g.sbl = sal0; g.sbj = sai0;
  std::swap(g.sci, g.scj);
}

/* hooks of the all-host tensor-path qgemm: slabs of rows arrive on the copy-in stream (events ev_in[p]), finished passes leave on
 * the copy-out stream */
/* Pageable host buffers (std::vector<Sleef_quad>, numpy: what callers of the reference API pass): a cudaMemcpyAsync from / to such
 * memory is staged by the driver at a fraction of the PCIe rate AND blocks the calling thread, which here is the thread that
 * enqueues the kernels.  So the pipelined qgemm hands the transfers to two feeder threads: the uploader runs the whole upload
 * sequence (paged_copy through ring 0 for pageable sources: four threads copy into page-locked slots while the copy engine moves
 * the previous ones), records the slab events and publishes a flag per slab, which the enqueueing thread waits for before it makes
 * a stream wait for that event (an event must be recorded before it is waited on); the downloader takes (event, range) jobs,
 * waits for the event and brings the rows home through ring 1.  8192^3 from numpy arrays: 275 ms -> see DESIGN.md §7. */
struct HostFeeder {
  std::vector<std::atomic<int>> up_done;          /* [P + 1]: slab p (P = the shared operand) is uploaded and its event recorded */
  std::atomic<int> err{(int)cudaSuccess};
  std::thread up, dn;
  struct Job { cudaEvent_t ev; char *host; const char *dev; size_t bytes; };
  std::mutex mu; std::condition_variable cv; std::deque<Job> jobs; bool closing = false;
  bool threaded_up = false, threaded_dn = false;
  explicit HostFeeder(int n) : up_done((size_t)n) { for (auto &f : up_done) f.store(0, std::memory_order_relaxed); }
  void wait_up(int idx) { while (!up_done[(size_t)idx].load(std::memory_order_acquire)) std::this_thread::yield(); }
  void fail_all(cudaError_t e) { err.store((int)e); for (auto &f : up_done) f.store(1, std::memory_order_release); }
  void push(cudaEvent_t ev, void *host, const void *dev, size_t bytes)
  {
    { std::lock_guard<std::mutex> lk(mu); jobs.push_back(Job{ev, (char *)host, (const char *)dev, bytes}); }
    cv.notify_one();
  }
  void download_loop(int device)
  {
    cudaSetDevice(device);
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return closing || !jobs.empty(); });
        if (jobs.empty()) return;
        j = jobs.front(); jobs.pop_front();
      }
      cudaError_t e = cudaEventSynchronize(j.ev);
      if (e == cudaSuccess) e = paged_copy(j.host, j.dev, j.bytes, false, 1);
      if (e != cudaSuccess) err.store((int)e);
    }
  }
  void finish()   /* every job handed over has been carried out, both threads are gone */
  {
    if (up.joinable()) up.join();
    if (dn.joinable()) { { std::lock_guard<std::mutex> lk(mu); closing = true; } cv.notify_all(); dn.join(); }
  }
  ~HostFeeder() { finish(); }
};

struct HostPipe {
  cudaEvent_t *ev_in; int P; int64_t blk;
  cudaStream_t ks, ds;
  EventSet *out_ev; int out_used;
  void *dC; void *hC; int64_t outer, inner, ldc;   /* C as (outer x inner) storage: slabs along `outer` are contiguous */
  cudaError_t err;
  HostFeeder *feed = nullptr;
};
static int host_rows_in(int64_t r0, int64_t rows, void *sA, void *sF, void *user)
{
  HostPipe *hp = (HostPipe *)user;
  int idx = (int)((r0 + rows - 1) / hp->blk);
  if (idx >= hp->P) idx = hp->P - 1;
  if (hp->feed) hp->feed->wait_up(idx);            /* the event is recorded by the uploader thread: not before this */
  cudaError_t e = cudaStreamWaitEvent((cudaStream_t)sA, hp->ev_in[idx], 0);
  if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)sF, hp->ev_in[idx], 0);
  if (e != cudaSuccess) { hp->err = e; return 1; }
  return 0;
}
static void host_rows_out(int64_t r0, int64_t rows, void *user)
{
  HostPipe *hp = (HostPipe *)user;
  if (hp->err != cudaSuccess || rows <= 0) return;
  if (!hp->out_ev->make(hp->out_used + 1)) { hp->err = cudaErrorMemoryAllocation; return; }
  cudaEvent_t ev = hp->out_ev->ev[hp->out_used++];
  cudaError_t e = cudaEventRecord(ev, hp->ks);            /* ks already waits for the reconstruction of these rows */
  const size_t off = (size_t)r0 * hp->ldc * 16, bytes = ((size_t)(rows - 1) * hp->ldc + hp->inner) * 16;
  if (e == cudaSuccess && hp->feed && hp->feed->threaded_dn) { hp->feed->push(ev, (char *)hp->hC + off, (char *)hp->dC + off, bytes); return; }
  if (e == cudaSuccess) e = cudaStreamWaitEvent(hp->ds, ev, 0);
  if (e == cudaSuccess) e = cudaMemcpyAsync((char *)hp->hC + off, (char *)hp->dC + off, bytes, cudaMemcpyDeviceToHost, hp->ds);
  if (e != cudaSuccess) hp->err = e;
}

/* `hooks`: the call comes from qb_gemm_dev, the only entry point that takes the row-pass callback, the streamed-B panels and
 * the peer outputs (the host paths below call this function per slab and must not fire them) */
static int gemm_dev_impl(char layout, char ta, char tb, int64_t m, int64_t n, int64_t k, q128 alpha, const void *dA,
                         int64_t lda, const void *dB, int64_t ldb, q128 beta, void *dC, int64_t ldc, cudaStream_t st, bool hooks = false)
{
  if (m < 0 || n < 0 || k < 0) return fail(QB_ERR_ARG, "qgemm: negative dimension");
  GemmArgs g = make_gemm_args(layout, ta, tb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc);
  const int mode = g_mode.load(), tp = g_tensor.load();
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  const bool streamed = hooks && g_bp_cb != nullptr;
  if (streamed && !(mode == QB_MODE_FAST && tp != 0)) return fail(QB_ERR_ARG, "qgemm: streamed B panels need the fast-mode tensor path");
  if (mode == QB_MODE_FAST && m > 0 && n > 0 && k > 0 && (tp == 2 || streamed || (tp == 1 && m >= 128 && n >= 128 && k >= 256))) {
    /* tensor-core path (exact int8 residues, qb_ozaki.cu); declines -> integer-limb kernel below */
    int used = 0;
    OzHooks h;
    if (hooks) {
      h.cb = (oz_pass_cb)g_pass_cb; h.cb_user = g_pass_user; h.min_passes = g_pass_min;
      h.bp = (oz_bpanel_cb)g_bp_cb; h.bp_user = g_bp_user; h.bp_cols = g_bp_cols; h.bstats = g_bp_stats;
      if (g_bp_cb) { h.bplanes_N = g_bp_planes_N; h.bplanes_W = g_bp_planes_W; }
      g.npeer = g_npeer;
      for (int q = 0; q < g.npeer; ++q) g.peerC[q] = (q128 *)g_peer[q];
    }
    g_peer_written = 0;
    cudaError_t oe = launch_gemm_ozaki(g, st, &used, h);
    if (oe != cudaSuccess) return fail(QB_ERR_CUDA, "qgemm tensor path", oe);
    if (used) { g_peer_written = oz_last_stats(false).peer_written; return QB_OK; }
    if (streamed) return fail(QB_ERR_ALLOC, "qgemm: the tensor path declined a streamed-B call (workspace)");
  }
  g_peer_written = 0;
  cudaError_t e = launch_gemm(g, mode, st);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemm kernel launch", e);
  if (hooks && g_pass_cb && m > 0) g_pass_cb(0, m, g_pass_user);   /* the integer-limb kernel produces all rows in one launch */
  return QB_OK;
}

} // namespace qb

using namespace qb;

extern "C" {

/* ------------------------------------------------------------------ housekeeping */
int qb_init(void)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  cudaDeviceProp pr;
  cudaError_t e = cudaGetDeviceProperties(&pr, g_curdev);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "cudaGetDeviceProperties", e);
  if (pr.major < 10) return fail(QB_ERR_CUDA, "qblas_b200 is built for sm_100a only; this device is older");
  return QB_OK;
}
const char *qb_last_error(void) { return t_err_msg; }
int qb_last_error_code(void) { return t_err_code; }
void qb_clear_error(void) { clear_error(); }
const char *qb_build_info(void) { return "qblas_b200 1.0.0 (sm_100a, integer-limb binary128; modes: reference-order, fast)"; }

void qb_set_mode(int mode) { g_mode.store(mode == QB_MODE_FAST ? QB_MODE_FAST : QB_MODE_REFERENCE); }
int qb_get_mode(void) { return g_mode.load(); }
void qb_set_tensor_path(int v) { g_tensor.store(v < 0 ? 0 : (v > 2 ? 2 : v)); }
int qb_get_tensor_path(void) { return g_tensor.load(); }
void qb_set_fast_variant(int v) { g_fastvar.store(v <= 0 ? 0 : (v >= 3 ? 3 : v)); }
int qb_get_fast_variant(void) { return g_fastvar.load(); }
void qb_set_ref_gemm_kernel(int v) { g_gemm_kernel.store(v != 0 ? 1 : 0); }
int qb_get_ref_gemm_kernel(void) { return g_gemm_kernel.load(); }
void qb_set_beta0_classes(int v) { g_beta0_classes.store(v ? 1 : 0); }
int qb_get_beta0_classes(void) { return g_beta0_classes.load(); }
void qb_set_gemm_pass_callback(qb_pass_cb cb, void *user, int min_passes)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  g_pass_cb = cb; g_pass_user = user; g_pass_min = min_passes < 1 ? 1 : min_passes;
}
void qb_set_gemm_b_panels(qb_bpanel_cb cb, void *user, int64_t panel_cols, const void *d_colstats)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  g_bp_cb = cb; g_bp_user = user; g_bp_cols = cb ? panel_cols : 0; g_bp_stats = cb ? (const int *)d_colstats : nullptr;
}
int qb_gemm_colstats_dev(char layout, char transb, int64_t k, int64_t n, const void *dB, int64_t ldb, void *d_colstats, void *stream)
{
  if (k < 0 || n < 0) return fail(QB_ERR_ARG, "qb_gemm_colstats_dev: negative dimension");
  if (k == 0 || n == 0) return QB_OK;
  const bool col = is_col(layout), tB = g_honor_trans.load() != 0 && is_trans(transb);
  const int64_t sbl = (col != tB) ? 1 : ldb, sbj = (col != tB) ? ldb : 1;
  cudaError_t e = launch_colstats((const q128 *)dB, n, k, sbj, sbl, (int *)d_colstats, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "column statistics launch", e);
  return QB_OK;
}
void qb_set_gemm_b_planes(int moduli, int window)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  g_bp_planes_N = moduli > 0 ? moduli : 0; g_bp_planes_W = moduli > 0 ? window : 0;
}
int qb_crt_plan(int span_a, int span_b, int64_t k, int *moduli, int *window_a, int *window_b)
{
  crt::host::Windows w;
  if (k <= 0 || !crt::host::plan_windows(span_a, span_b, k, oz_get_window(), w)) return -1;
  if (moduli) *moduli = w.N;
  if (window_a) *window_a = w.WA;
  if (window_b) *window_b = w.WB;
  return (w.truncA ? 1 : 0) | (w.truncB ? 2 : 0);
}
int qb_crt_residues_dev(char layout, char transb, int64_t k, int64_t n, const void *dB, int64_t ldb, const void *d_emax, int window, int moduli,
                        void *d_planes, int64_t plane_stride, void *stream)
{
  if (k <= 0 || n <= 0 || moduli < 1 || moduli > crt::NM || window < 1 || window > crt::WMAX) return fail(QB_ERR_ARG, "qb_crt_residues_dev: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  cudaError_t e = oz_prepare_device();
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qb_crt_residues_dev", e);
  const bool col = is_col(layout), tB = g_honor_trans.load() != 0 && is_trans(transb);
  const int64_t sbl = (col != tB) ? 1 : ldb, sbj = (col != tB) ? ldb : 1;
  const int64_t Kp = (k + 127) / 128 * 128;
  launch_crt_residues((const q128 *)dB, n, k, sbj, sbl, (const int *)d_emax, window, moduli, Kp, (int8_t *)d_planes, (cudaStream_t)stream, plane_stride);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "residue kernel launch", e);
  return QB_OK;
}
void qb_set_tensor_window(int bits)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  oz_set_window(bits);
}
int qb_get_tensor_window(void) { return oz_get_window(); }
void qb_set_tensor_unit(int64_t rows, int64_t cols)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  oz_set_unit(rows, cols);
}
void qb_get_tensor_unit(int64_t *rows, int64_t *cols) { oz_get_unit(rows, cols); }
void qb_set_tensor_ramp(int64_t rows, int64_t cols)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  oz_set_ramp(rows, cols);
}
void qb_get_tensor_ramp(int64_t *rows, int64_t *cols) { oz_get_ramp(rows, cols); }
void qb_set_tensor_workspace_limit(size_t bytes)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  oz_set_ws_limit(bytes);
}
size_t qb_get_tensor_workspace_limit(void) { return oz_get_ws_limit(); }
void qb_set_host_slabs(int slabs) { g_host_slabs.store(slabs < 1 ? 1 : (slabs > 16 ? 16 : slabs)); }
int qb_get_host_slabs(void) { return g_host_slabs.load(); }
void qb_set_tensor_pass_shape(int shape)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  oz_set_pass_shape(shape);
}
int qb_get_tensor_pass_shape(void) { return oz_get_pass_shape(); }
int qb_crt_pass_rows(int64_t m, int64_t cap, int shape, int64_t *out, int max_out)
{
  const std::vector<int64_t> rows = oz_crt_pass_rows(m, cap, shape);
  for (int i = 0; i < (int)rows.size() && i < max_out; ++i) out[i] = rows[i];
  return (int)rows.size();
}

/* ---- peer memory (fused gather of the row-sharded qgemm) ---- */
void *qb_peer_alloc(size_t bytes)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (ensure_device()) return nullptr;
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) { fail(QB_ERR_ALLOC, "qb_peer_alloc", e); return nullptr; }
  return p;
}
void qb_peer_free(void *p) { if (p) cudaFree(p); }
int qb_peer_export(void *p, void *handle64)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "cudaIpcGetMemHandle", e);
  memcpy(handle64, &h, 64);
  return QB_OK;
}
void *qb_peer_open(const void *handle64)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (ensure_device()) return nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { fail(QB_ERR_CUDA, "cudaIpcOpenMemHandle", e); return nullptr; }
  return p;
}
int qb_peer_close(void *p)
{
  cudaError_t e = cudaIpcCloseMemHandle(p);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "cudaIpcCloseMemHandle", e);
  return QB_OK;
}
int qb_set_gemm_peer_outputs(int count, void *const *peer_C)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (count < 0 || count > QB_MAX_PEERS || (count > 0 && !peer_C)) return fail(QB_ERR_ARG, "qb_set_gemm_peer_outputs: 0 <= count <= 8");
  g_npeer = count;
  for (int q = 0; q < count; ++q) g_peer[q] = peer_C[q];
  return QB_OK;
}
int qb_get_gemm_peer_written(void) { return g_peer_written; }
void qb_oz_last_stats(int64_t *out16)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  const OzStats s = oz_last_stats();
  out16[0] = s.pairs; out16[1] = s.WA; out16[2] = s.WB; out16[3] = s.WA_nat; out16[4] = s.WB_nat; out16[5] = s.truncated; out16[6] = s.flagged;
  out16[7] = s.row_passes; out16[8] = s.panels; out16[9] = s.units; out16[10] = s.nchunks; out16[11] = s.Kp; out16[12] = s.ws_bytes;
  out16[13] = s.peer_written; out16[14] = 0; out16[15] = 0;
}
double qb_oz_last_mma_ms(int *launches)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  return oz_last_mma_ms(launches);
}
int qb_oz_last_mma_timeline(double *out, int max_pairs)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  return oz_last_mma_timeline(out, max_pairs);
}
int qb_oz_i8gemm_dev(const void *dPlanesA, const void *dPlanesB, int SA, int SB, int64_t m, int64_t n, int64_t Kp, int64_t kb_begin,
                     int64_t nkb, void *dD, int64_t Mp, int64_t Np, void *stream)
{
  if (SA < 1 || SB < 1 || SA > QB_OZ_MAX_SLICES || SB > QB_OZ_MAX_SLICES || Kp % 128 || Mp % 128 || Np % 256 || Mp < m || Np < n)
    return fail(QB_ERR_ARG, "qb_oz_i8gemm_dev: bad geometry");
  cudaError_t e = launch_oz_mma((const int8_t *)dPlanesA, (const int8_t *)dPlanesB, SA, SB, m, n, Kp, (int)kb_begin, (int)nkb, (int32_t *)dD, Mp,
                                Np, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "int8 diagonal GEMM launch", e);
  return QB_OK;
}
void qb_set_kc(int kc) { g_kc.store(kc > 0 ? kc : 126); }
int qb_get_kc(void) { return g_kc.load(); }
void qb_set_honor_trans(int on) { g_honor_trans.store(on ? 1 : 0); }
int qb_get_honor_trans(void) { return g_honor_trans.load(); }
int64_t qb_launch_count(void) { return g_launches.load(); }

/* QuadBLAS::aligned_alloc / aligned_free (memory/allocation.hpp:18-41): page-locked host memory so the
 * staging copies run at full rate; plain 32-byte aligned memory when pinning is refused. */
static std::mutex g_pin_mu;
static std::unordered_set<void *> g_pinned;
void *qb_host_alloc(size_t bytes)
{
  if (bytes == 0) bytes = 16;
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess && p) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pinned.insert(p);
    return p;
  }
  cudaGetLastError();
  if (posix_memalign(&p, 32, bytes) != 0) return nullptr;
  return p;
}
void qb_host_free(void *p)
{
  if (!p) return;
  bool pinned;
  { std::lock_guard<std::mutex> lk(g_pin_mu); pinned = g_pinned.erase(p) != 0; }
  if (pinned) cudaFreeHost(p); else free(p);
}

void quadblas_set_num_threads(int num_threads) { g_threads.store(num_threads); }
int quadblas_get_num_threads(void) { return num_threads(); }
const char *quadblas_get_version(void) { return "QuadBLAS 1.0.0 - High Performance Quad Precision BLAS"; }
int quadblas_is_aligned(const void *ptr) { return ((uintptr_t)ptr % 32) == 0; }

qb_quad qb_from_double(double d)
{
  uint64_t b; memcpy(&b, &d, 8);
  q128 q = q_from_double_bits(b);
  qb_quad r; r.lo = q.lo; r.hi = q.hi; return r;
}
double qb_to_double(qb_quad v)
{
  q128 q; q.lo = v.lo; q.hi = v.hi;
  uint64_t b = q_to_double_bits(q);
  double d; memcpy(&d, &b, 8); return d;
}

/* ------------------------------------------------------------------ device-pointer API */
int qb_gemm_dev(char layout, char transa, char transb, int64_t m, int64_t n, int64_t k, const qb_quad *alpha,
                const void *dA, int64_t lda, const void *dB, int64_t ldb, const qb_quad *beta, void *dC,
                int64_t ldc, void *stream)
{
  return gemm_dev_impl(layout, transa, transb, m, n, k, toq(alpha), dA, lda, dB, ldb, toq(beta), dC, ldc, (cudaStream_t)stream, true);
}

static int gemv_dev_impl(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda, const void *dx,
                         int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream, int64_t m_plan)
{
  if (m < 0 || n < 0) return fail(QB_ERR_ARG, "qgemv: negative dimension");
  GemvArgs g;
  g.m = m; g.n = n; g.alpha = toq(alpha); g.beta = toq(beta);
  g.A = (const q128 *)dA; g.lda = lda; g.col_major = is_col(layout);
  g.x = (const q128 *)dx; g.incx = incx; g.y = (q128 *)dy; g.incy = incy;
  g.work = nullptr; g.work_elems = 0; g.m_plan = m_plan;
  const int mode = g_mode.load();
  const int64_t need = gemv_work_elems(m, n, g.col_major, mode, m_plan);
  std::unique_lock<std::recursive_mutex> lk(g_mu, std::defer_lock);
  if (need > 0) { /* the shared scratch is held until the launch is queued (stream order protects its reuse) */
    lk.lock();
    int rc = ensure_device();
    if (rc) return rc;
    rc = ensure_work(need);
    if (rc) return rc;
    g.work = S().work; g.work_elems = S().work_elems;
  }
  cudaError_t e = launch_gemv(g, mode, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemv kernel launch", e);
  g_last_sliced.flags = need > 0 && mode != 0 ? gemv_sliced_rowflags(g.work, m, n, g.col_major, m_plan) : nullptr;
  g_last_sliced.m = m; g_last_sliced.dev = g_curdev;
  return QB_OK;
}

int qb_gemv_dev(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda, const void *dx,
                int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream)
{
  return gemv_dev_impl(layout, m, n, alpha, dA, lda, dx, incx, beta, dy, incy, stream, 0);
}

int qb_gemv_rows_dev(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *dA, int64_t lda, const void *dx,
                     int64_t incx, const qb_quad *beta, void *dy, int64_t incy, void *stream, int64_t m_total)
{
  if (m_total < m) return fail(QB_ERR_ARG, "qgemv: m_total smaller than the block");
  return gemv_dev_impl(layout, m, n, alpha, dA, lda, dx, incx, beta, dy, incy, stream, m_total);
}

/* rows of the last qb_gemv_dev on this thread's device that the sliced FP64 kernel declined and the window kernel recomputed
 * (-1: that call did not take the sliced path).  Synchronises the device: diagnostics / tests only. */
int64_t qb_gemv_last_declined(void)
{
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  if (g_last_sliced.flags == nullptr) return -1;
  std::vector<uint8_t> h((size_t)g_last_sliced.m);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(h.data(), g_last_sliced.flags, h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  int64_t c = 0;
  for (uint8_t f : h) c += f != 0;
  return c;
}

/* force_T > 0: reference order with that many chunks whatever the mode (qb_dot_kernel: T = 1 is dot_kernel_vectorized) */
static int dot_dev_impl(int64_t n, const void *dx, int64_t incx, const void *dy, int64_t incy, int do_sqrt,
                        void *d_result, cudaStream_t st, int force_T = 0)
{
  if (n < 0) return fail(QB_ERR_ARG, "qdot: negative n");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  const int mode = force_T > 0 ? (int)QB_MODE_REFERENCE : g_mode.load();
  const int T = force_T > 0 ? force_T : num_threads();
  rc = ensure_work(dot_work_elems(n, T, mode));
  if (rc) return rc;
  DotArgs g;
  g.n = n; g.x = (const q128 *)dx; g.incx = incx; g.y = (const q128 *)dy; g.incy = incy;
  g.T = T; g.do_sqrt = do_sqrt; g.result = (q128 *)d_result; g.work = S().work; g.work_elems = S().work_elems;
  g.ticket = (unsigned *)(S().result + 2);
  g.only_if = (unsigned *)(S().result + 3);
  cudaError_t e = launch_dot(g, mode, st);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qdot kernel launch", e);
  return QB_OK;
}

int qb_dot_dev(int64_t n, const void *dx, int64_t incx, const void *dy, int64_t incy, void *d_result, void *stream)
{ return dot_dev_impl(n, dx, incx, dy, incy, 0, d_result, (cudaStream_t)stream); }
int qb_nrm2_dev(int64_t n, const void *dx, int64_t incx, void *d_result, void *stream)
{ return dot_dev_impl(n, dx, incx, dx, incx, 1, d_result, (cudaStream_t)stream); }

int qb_axpy_dev(int64_t n, const qb_quad *alpha, const void *dx, int64_t incx, void *dy, int64_t incy, void *stream)
{
  cudaError_t e = launch_axpy(n, toq(alpha), (const q128 *)dx, incx, (q128 *)dy, incy, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qaxpy kernel launch", e);
  return QB_OK;
}

int qb_dot_partials_dev(int64_t n_local, const void *dx, int64_t incx, const void *dy, int64_t incy, int64_t chunk,
                        int64_t nchunks, void *d_partials, void *stream)
{
  if (n_local < 0 || nchunks < 0 || chunk < 0) return fail(QB_ERR_ARG, "qdot partials: negative argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  rc = ensure_work(3 * nchunks + 4);
  if (rc) return rc;
  DotArgs g;
  g.n = n_local; g.x = (const q128 *)dx; g.incx = incx; g.y = (const q128 *)dy; g.incy = incy;
  g.T = (int)nchunks; g.do_sqrt = 0; g.result = nullptr; g.work = S().work; g.work_elems = S().work_elems;
  cudaError_t e = launch_dot_partials(g, chunk, (int)nchunks, (q128 *)d_partials, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qdot partials launch", e);
  return QB_OK;
}

int qb_fold_partials_dev(int64_t count, const void *d_partials, int do_sqrt, void *d_result, void *stream)
{
  cudaError_t e = launch_fold(count, (const q128 *)d_partials, do_sqrt, (q128 *)d_result, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "fold kernel launch", e);
  return QB_OK;
}

int qb_elementwise_dev(int op, int64_t n, const void *da, const void *db, const void *dc, void *dout, void *stream)
{
  cudaError_t e = launch_elementwise(op, n, (const q128 *)da, (const q128 *)db, (const q128 *)dc, (q128 *)dout, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "elementwise kernel launch", e);
  return QB_OK;
}

int qb_fma_microbench_dev(int variant, int blocks, int threads, int iters, void *d_sink, int64_t *n_fma, void *stream)
{
  cudaError_t e = launch_fma_microbench(variant, blocks, threads, iters, (q128 *)d_sink, n_fma, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "microbench launch", e);
  return QB_OK;
}

/* ------------------------------------------------------------------ host-or-device synchronous API */
int qb_gemm(char layout, char transa, char transb, int64_t m, int64_t n, int64_t k, const qb_quad *alpha, const void *A,
            int64_t lda, const void *B, int64_t ldb, const qb_quad *beta, void *C, int64_t ldc)
{
  if (m < 0 || n < 0 || k < 0) return fail(QB_ERR_ARG, "qgemm: negative dimension");
  if (m == 0 || n == 0 || k == 0) return QB_OK; /* level3.hpp:221 */
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  const bool col = is_col(layout);
  const bool honor = g_honor_trans.load() != 0;
  const bool tA = honor && is_trans(transa), tB = honor && is_trans(transb);
  /* footprints: A is (col != tA) ? k-major : m-major */
  const size_t a_bytes = (col != tA) ? mat_bytes(k, m, lda) : mat_bytes(m, k, lda);
  const size_t b_bytes = (col != tB) ? mat_bytes(n, k, ldb) : mat_bytes(k, n, ldb);
  const size_t c_bytes = col ? mat_bytes(n, m, ldc) : mat_bytes(m, n, ldc);
  const void *dA, *dB, *dC; bool sa, sb, sc;
  /* Large all-host problems: pipeline the transfers against the compute.  The operand that every block
   * needs (B for row-major, A for col-major) goes first; then C is cut into P contiguous slabs (row
   * blocks for row-major, column blocks for col-major) and for each slab the matching block of the
   * other operand and the slab of C_in are uploaded on a copy stream while earlier slabs compute, and
   * each finished slab is downloaded on a third stream while the next one computes.  Every slab is an
   * ordinary qgemm on device pointers, so the arithmetic (and, in reference-order mode, every bit) is
   * the same as the unpipelined call. */
  const int64_t split = col ? n : m;
  const bool all_host = !is_device_ptr(A) && !is_device_ptr(B) && !is_device_ptr(C);
  if (all_host && split >= 1024 && a_bytes + b_bytes + c_bytes >= ((size_t)64 << 20)) {
    if ((rc = stage_in(0, A, a_bytes, false, &dA, &sa))) return rc;
    if ((rc = stage_in(1, B, b_bytes, false, &dB, &sb))) return rc;
    if ((rc = stage_in(2, C, c_bytes, false, &dC, &sc))) return rc;
    if ((rc = ensure_streams())) return rc;
    const cudaStream_t cs = S().cs, ks = S().ks, ds = S().ds;
    constexpr int PMAX = 16;
    const int P = std::min(PMAX, std::max(1, g_host_slabs.load()));   /* qb_set_host_slabs */
    EventSet evs;
    if (!evs.make(2 * P + 1)) return fail(QB_ERR_CUDA, "qgemm: event creation", cudaGetLastError());
    cudaEvent_t *ev_in = evs.ev.data(), *ev_done = evs.ev.data() + P + 1;
    /* block [c0, c0+cnt) of a matrix X(outer, inner) with leading dimension ld, taken along `along_outer`:
     * a contiguous slab when taken along the strided direction, a 2-D copy otherwise */
    auto xfer = [&](void *dev, void *host, bool to_dev, bool along_outer, int64_t outer, int64_t inner, int64_t ld, int64_t c0, int64_t cnt,
                    cudaStream_t st) -> cudaError_t {
      if (cnt <= 0) return cudaSuccess;
      char *d = (char *)dev, *h = (char *)host;
      if (along_outer) {
        /* whole rows including the padding up to the next block's first row (a download may cut the rows differently than the
         * uploads did: the padding in between must hold the caller's bytes); the last row of the matrix ends at its last element */
        const size_t off = (size_t)c0 * ld * 16, bytes = (c0 + cnt < outer ? (size_t)cnt * ld : (size_t)(cnt - 1) * ld + inner) * 16;
        return to_dev ? cudaMemcpyAsync(d + off, h + off, bytes, cudaMemcpyHostToDevice, st) : cudaMemcpyAsync(h + off, d + off, bytes, cudaMemcpyDeviceToHost, st);
      }
      const size_t off = (size_t)c0 * 16;
      return to_dev ? cudaMemcpy2DAsync(d + off, (size_t)ld * 16, h + off, (size_t)ld * 16, (size_t)cnt * 16, (size_t)outer, cudaMemcpyHostToDevice, st)
                    : cudaMemcpy2DAsync(h + off, (size_t)ld * 16, d + off, (size_t)ld * 16, (size_t)cnt * 16, (size_t)outer, cudaMemcpyDeviceToHost, st);
    };
    cudaError_t e = cudaSuccess;
    /* beta = +-0 and contiguous rows of C: classify C_in on the host instead of uploading it (see CodeScan) */
    const q128 bq = toq(beta);
    const int64_t c_outer = col ? n : m, c_inner = col ? m : n;
    std::unique_ptr<CodeScan> scan;
    if ((bq.hi & 0x7fffffffffffffffull) == 0 && bq.lo == 0 && ldc == c_inner && g_beta0_classes.load() != 0) {
      const size_t need = (size_t)c_outer * (size_t)c_inner;
      Scratch &sc_ = S();
      bool ok = true;
      if (sc_.codes_bytes < need) {
        if (sc_.codes_h) cudaFreeHost(sc_.codes_h);
        if (sc_.codes_d) cudaFree(sc_.codes_d);
        sc_.codes_h = nullptr; sc_.codes_d = nullptr; sc_.codes_bytes = 0;
        ok = cudaHostAlloc((void **)&sc_.codes_h, need, cudaHostAllocDefault) == cudaSuccess && cudaMalloc((void **)&sc_.codes_d, need) == cudaSuccess;
        if (ok) sc_.codes_bytes = need;
        else { if (sc_.codes_h) cudaFreeHost(sc_.codes_h); if (sc_.codes_d) cudaFree(sc_.codes_d); sc_.codes_h = nullptr; sc_.codes_d = nullptr; cudaGetLastError(); }
      }
      if (ok) {
        try {
          const int64_t chunk = 1 << 16, nch = ((int64_t)need + chunk - 1) / chunk;
          scan.reset(new CodeScan(nch));
          scan->C = (const q128 *)C; scan->codes = sc_.codes_h; scan->count = (int64_t)need; scan->chunk = chunk; scan->nchunks = nch;
          const unsigned hw = std::thread::hardware_concurrency();
          const int T = (int)std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
          for (int t = 0; t < T; ++t) scan->th.emplace_back([sp = scan.get()] { sp->work(); });
        } catch (...) {
          if (scan && scan->th.empty()) scan.reset();            /* no thread at all: upload C as before; some threads: they finish the scan */
        }
      }
    }
    /* storage shapes (outer x inner): A is k x m when (col != tA) else m x k; B is n x k when (col != tB) else k x n */
    const bool a_m_outer = !(col != tA), b_n_outer = (col != tB);
    const int64_t blk = ((split + P - 1) / P + 127) / 128 * 128;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    HostFeeder feed(P + 1);
    const bool page_in = is_pageable(A) || is_pageable(B) || (!scan && is_pageable(C));
    const bool page_out = is_pageable(C);
    /* an upload of a slab taken along the strided direction of a pageable matrix goes through the page-locked ring (synchronous:
     * that is why it runs on the uploader thread); everything else is an asynchronous copy as before */
    auto up_xfer = [&](void *dev, const void *host, bool along_outer, int64_t outer, int64_t inner, int64_t ld, int64_t c0, int64_t cnt) -> cudaError_t {
      if (cnt > 0 && along_outer && page_in && is_pageable(host)) {
        const size_t off = (size_t)c0 * ld * 16, bytes = (c0 + cnt < outer ? (size_t)cnt * ld : (size_t)(cnt - 1) * ld + inner) * 16;
        return paged_copy((char *)dev + off, (const char *)host + off, bytes, true, 0);
      }
      return xfer(dev, const_cast<void *>(host), true, along_outer, outer, inner, ld, c0, cnt, cs);
    };
    auto upload_all = [&]() {
      cudaError_t ue = cudaSuccess;
      if (page_in) ue = cudaSetDevice(cur_dev);
      if (ue == cudaSuccess) {                                   /* shared operand first */
        const void *src = !col ? B : A; void *dst = !col ? (void *)dB : (void *)dA; const size_t bytes = !col ? b_bytes : a_bytes;
        ue = (page_in && is_pageable(src)) ? paged_copy(dst, src, bytes, true, 0) : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs);
      }
      if (ue == cudaSuccess) ue = cudaEventRecord(ev_in[P], cs);
      if (ue != cudaSuccess) { feed.fail_all(ue); return; }
      feed.up_done[(size_t)P].store(1, std::memory_order_release);
      for (int p = 0; p < P; ++p) {
        const int64_t c0 = (int64_t)p * blk, cnt = std::min(blk, split - c0);
        if (cnt > 0) {
          if (!col) ue = up_xfer((void *)dA, A, a_m_outer, a_m_outer ? m : k, a_m_outer ? k : m, lda, c0, cnt);
          else ue = up_xfer((void *)dB, B, b_n_outer, b_n_outer ? n : k, b_n_outer ? k : n, ldb, c0, cnt);
          if (ue == cudaSuccess && !scan) ue = up_xfer((void *)dC, C, true, col ? n : m, col ? m : n, ldc, c0, cnt);   /* beta*C is always read */
          if (ue == cudaSuccess && scan) {   /* ... through its classes: one byte per element, expanded to +1 / -1 / NaN on the device */
            const int64_t lo = c0 * c_inner, cntel = cnt * c_inner;
            scan->wait_elems(lo, lo + cntel);
            ue = cudaMemcpyAsync(S().codes_d + lo, S().codes_h + lo, (size_t)cntel, cudaMemcpyHostToDevice, cs);
            if (ue == cudaSuccess) {
              k_c_standin<<<(unsigned)std::min<int64_t>((cntel + 255) / 256, 148 * 8), 256, 0, cs>>>(S().codes_d + lo, (q128 *)dC + lo, cntel);
              count_launch();
              ue = cudaGetLastError();
            }
          }
        }
        if (ue == cudaSuccess) ue = cudaEventRecord(ev_in[p], cs);
        if (ue != cudaSuccess) { feed.fail_all(ue); return; }
        feed.up_done[(size_t)p].store(1, std::memory_order_release);
      }
    };
    if (page_in) {
      try { feed.up = std::thread(upload_all); feed.threaded_up = true; } catch (...) { feed.threaded_up = false; }
    }
    if (!feed.threaded_up) upload_all();
    if (page_out) {
      try { feed.dn = std::thread([&feed, cur_dev] { feed.download_loop(cur_dev); }); feed.threaded_dn = true; } catch (...) { feed.threaded_dn = false; }
    }
    feed.wait_up(P);
    e = (cudaError_t)feed.err.load();
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ks, ev_in[P], 0);
    /* fast mode: ONE tensor-path call whose row passes wait for their slabs (passes outer: the residue planes of the shared
     * operand are computed once and stay resident) and whose finished passes are downloaded while the next ones compute */
    bool streamed = false;
    const int mode = g_mode.load(), tp = g_tensor.load();
    if (e == cudaSuccess && mode == QB_MODE_FAST && (tp == 2 || (tp == 1 && m >= 128 && n >= 128 && k >= 256))) {
      GemmArgs g = make_gemm_args(layout, transa, transb, m, n, k, toq(alpha), dA, lda, dB, ldb, toq(beta), (void *)dC, ldc);
      if (col) swap_roles(g);
      EventSet out_ev;
      HostPipe hp;
      hp.ev_in = ev_in; hp.P = P; hp.blk = blk; hp.ks = ks; hp.ds = ds; hp.out_ev = &out_ev; hp.out_used = 0;
      hp.dC = (void *)dC; hp.hC = C; hp.outer = col ? n : m; hp.inner = col ? m : n; hp.ldc = ldc; hp.err = cudaSuccess; hp.feed = &feed;
      OzHooks h;
      h.order = 1; h.rows_in = host_rows_in; h.rows_user = &hp; h.cb = host_rows_out; h.cb_user = &hp;
      h.min_passes = (int)std::min<int64_t>(4096, (split + 1023) / 1024);     /* passes of <= 1024 rows: a short tail after the last upload */
      int used = 0;
      const cudaError_t oe = launch_gemm_ozaki(g, ks, &used, h);
      if (oe != cudaSuccess || hp.err != cudaSuccess) {
        feed.finish();
        cudaStreamSynchronize(cs); cudaStreamSynchronize(ks); cudaStreamSynchronize(ds);
        return fail(QB_ERR_CUDA, "qgemm pipelined host path (tensor)", oe != cudaSuccess ? oe : hp.err);
      }
      streamed = used != 0;
      if (streamed) { feed.finish(); cudaStreamSynchronize(cs); cudaStreamSynchronize(ks); cudaStreamSynchronize(ds); }   /* out_ev is destroyed with this scope */
    }
    for (int p = 0; p < P && e == cudaSuccess && rc == QB_OK && !streamed; ++p) {
      const int64_t c0 = (int64_t)p * blk, cnt = std::min(blk, split - c0);
      if (cnt <= 0) break;
      feed.wait_up(p);
      e = cudaStreamWaitEvent(ks, ev_in[p], 0);
      if (e != cudaSuccess) break;
      /* element offsets of the block inside A / B / C (same strides as gemm_dev_impl derives) */
      const int64_t a_off = !col ? c0 * (a_m_outer ? lda : 1) : 0;
      const int64_t b_off = col ? c0 * (b_n_outer ? ldb : 1) : 0;
      const int64_t c_off = c0 * ldc;
      rc = gemm_dev_impl(layout, transa, transb, col ? m : cnt, col ? cnt : n, k, toq(alpha), (const q128 *)dA + a_off, lda, (const q128 *)dB + b_off, ldb,
                         toq(beta), (q128 *)dC + c_off, ldc, ks);
      if (rc) break;
      e = cudaEventRecord(ev_done[p], ks);
      if (e == cudaSuccess && feed.threaded_dn) {
        const int64_t c_out = col ? n : m, c_in = col ? m : n;
        const size_t off = (size_t)c0 * ldc * 16, bytes = (c0 + cnt < c_out ? (size_t)cnt * ldc : (size_t)(cnt - 1) * ldc + c_in) * 16;
        feed.push(ev_done[p], (char *)C + off, (const char *)dC + off, bytes);
        continue;
      }
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ds, ev_done[p], 0);
      if (e == cudaSuccess) e = xfer((void *)dC, C, false, true, col ? n : m, col ? m : n, ldc, c0, cnt, ds);
    }
    feed.finish();
    if (e == cudaSuccess) e = (cudaError_t)feed.err.load();
    cudaError_t e2 = cudaStreamSynchronize(cs), e3 = cudaStreamSynchronize(ks), e4 = cudaStreamSynchronize(ds);
    if (rc) return rc;
    if (e == cudaSuccess) e = e2 != cudaSuccess ? e2 : (e3 != cudaSuccess ? e3 : e4);
    if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemm pipelined host path", e);
    return QB_OK;
  }
  if ((rc = stage_in(0, A, a_bytes, true, &dA, &sa))) return rc;
  if ((rc = stage_in(1, B, b_bytes, true, &dB, &sb))) return rc;
  if ((rc = stage_in(2, C, c_bytes, true, &dC, &sc))) return rc; /* beta*C is always read (level3.hpp:107) */
  rc = gemm_dev_impl(layout, transa, transb, m, n, k, toq(alpha), dA, lda, dB, ldb, toq(beta), (void *)dC, ldc, 0);
  if (rc) return rc;
  cudaError_t e;
  if (sc) { e = cudaDeviceSynchronize(); if (e == cudaSuccess) e = paged_copy(C, dC, c_bytes, false); } else e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemm completion", e);
  return QB_OK;
}

int qb_gemv(char layout, int64_t m, int64_t n, const qb_quad *alpha, const void *A, int64_t lda, const void *x,
            int64_t incx, const qb_quad *beta, void *y, int64_t incy)
{
  if (m < 0 || n < 0) return fail(QB_ERR_ARG, "qgemv: negative dimension");
  if (m == 0 || n == 0) return QB_OK;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  const bool col = is_col(layout);
  const size_t a_bytes = col ? mat_bytes(n, m, lda) : mat_bytes(m, n, lda);
  const size_t x_bytes = vec_bytes(n, incx), y_bytes = vec_bytes(m, incy);
  const void *dA, *dx, *dy; bool sa, sx, sy;
  /* Large host A (16 GiB at 32768^2): the upload is the whole cost, so A goes up in row slabs (rows of y) on a copy stream
   * while the kernels of the earlier slabs run: contiguous blocks for row-major, 2-D copies of a row range of every column
   * for col-major.  Each slab is an ordinary qgemv on its rows: every y_i sees exactly the arithmetic of the unsplit call. */
  if (!is_device_ptr(A) && m >= 1024 && a_bytes >= ((size_t)64 << 20)) {
    if ((rc = stage_in(0, A, a_bytes, false, &dA, &sa))) return rc;
    if ((rc = stage_in(1, x, x_bytes, true, &dx, &sx))) return rc;
    if ((rc = stage_in(2, y, y_bytes, true, &dy, &sy))) return rc;
    if ((rc = ensure_streams())) return rc;
    const cudaStream_t cs = S().cs, ks = S().ks;
    constexpr int PMAX = 16;
    const int P = std::min(PMAX, std::max(1, g_host_slabs.load()));
    EventSet evs;
    if (!evs.make(P)) return fail(QB_ERR_CUDA, "qgemv: event creation", cudaGetLastError());
    const int64_t blk = ((m + P - 1) / P + 31) / 32 * 32;
    const bool a_pageable = is_pageable(A);
    cudaError_t e = cudaSuccess;
    for (int p = 0; p < P && e == cudaSuccess; ++p) {
      const int64_t r0 = (int64_t)p * blk, cnt = std::min(blk, m - r0);
      if (cnt <= 0) break;
      if (!col) {
        const size_t off = (size_t)r0 * lda * 16, bytes = ((size_t)(cnt - 1) * lda + n) * 16;
        /* pageable A: through the page-locked ring (four copy threads; synchronous, but the kernels of the earlier slabs are already
         * queued, so the device multiplies slab p - 1 while this thread copies slab p) */
        if (a_pageable) e = paged_copy((char *)dA + off, (const char *)A + off, bytes, true, 0);
        else e = cudaMemcpyAsync((char *)dA + off, (const char *)A + off, bytes, cudaMemcpyHostToDevice, cs);
      } else {
        const size_t off = (size_t)r0 * 16;
        e = cudaMemcpy2DAsync((char *)dA + off, (size_t)lda * 16, (const char *)A + off, (size_t)lda * 16, (size_t)cnt * 16, (size_t)n, cudaMemcpyHostToDevice, cs);
      }
      if (e == cudaSuccess) e = cudaEventRecord(evs.ev[p], cs);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(ks, evs.ev[p], 0);
      if (e != cudaSuccess) break;
      const q128 *As = (const q128 *)dA + (col ? r0 : r0 * lda);
      rc = gemv_dev_impl(layout, cnt, n, alpha, As, lda, dx, incx, beta, (q128 *)dy + r0 * incy, incy, ks, m);   /* planned as the whole call: same bits */
      if (rc) break;
    }
    const cudaError_t e2 = cudaStreamSynchronize(cs), e3 = cudaStreamSynchronize(ks);
    if (rc) return rc;
    if (e == cudaSuccess) e = e2 != cudaSuccess ? e2 : e3;
    if (e == cudaSuccess && sy) e = cudaMemcpy(y, dy, y_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemv pipelined host path", e);
    return QB_OK;
  }
  if ((rc = stage_in(0, A, a_bytes, true, &dA, &sa))) return rc;
  if ((rc = stage_in(1, x, x_bytes, true, &dx, &sx))) return rc;
  if ((rc = stage_in(2, y, y_bytes, true, &dy, &sy))) return rc;
  rc = qb_gemv_dev(layout, m, n, alpha, dA, lda, dx, incx, beta, (void *)dy, incy, 0);
  if (rc) return rc;
  cudaError_t e;
  if (sy) e = cudaMemcpy(y, dy, y_bytes, cudaMemcpyDeviceToHost); else e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qgemv completion", e);
  return QB_OK;
}

static int dot_host_impl(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, int do_sqrt, qb_quad *result, int force_T = 0)
{
  if (n < 0) return fail(QB_ERR_ARG, "qdot: negative n");
  const void *dx, *dy; bool sx = false, sy = false;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  if ((rc = stage_in(0, x, vec_bytes(n, incx), true, &dx, &sx))) return rc;
  if (y == x && incy == incx) { dy = dx; }   /* same vector: one staged copy serves both (different strides read different extents) */
  else if ((rc = stage_in(1, y, vec_bytes(n, incy), true, &dy, &sy))) return rc;
  rc = dot_dev_impl(n, dx, incx, dy, incy, do_sqrt, S().result, 0, force_T);
  if (rc) return rc;
  q128 r;
  cudaError_t e = cudaMemcpy(&r, S().result, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qdot result copy", e);
  result->lo = r.lo; result->hi = r.hi;
  return QB_OK;
}

int qb_dot(int64_t n, const void *x, int64_t incx, const void *y, int64_t incy, qb_quad *result)
{ return dot_host_impl(n, x, incx, y, incy, 0, result); }
int qb_nrm2(int64_t n, const void *x, int64_t incx, qb_quad *result)
{ return dot_host_impl(n, x, incx, x, incx, 1, result); }
/* QuadBLAS::dot_kernel_vectorized (level1.hpp:14-35): the two-lane kernel over contiguous data, whatever the mode and the thread
 * count: even / odd chains from +0, add(lane0, lane1), odd tail.  (dot with ONE chunk is that kernel followed by add(+0, r),
 * and r is never -0, so the bits are the kernel's.) */
int qb_dot_kernel(int64_t n, const void *x, const void *y, qb_quad *result)
{ return dot_host_impl(n, x, 1, y, 1, 0, result, 1); }

int qb_axpy(int64_t n, const qb_quad *alpha, const void *x, int64_t incx, void *y, int64_t incy)
{
  if (n <= 0) return QB_OK; /* c_interface.hpp:49 */
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int rc = ensure_device();
  if (rc) return rc;
  const void *dx, *dy; bool sx, sy;
  if ((rc = stage_in(0, x, vec_bytes(n, incx), true, &dx, &sx))) return rc;
  if ((rc = stage_in(1, y, vec_bytes(n, incy), true, &dy, &sy))) return rc;
  rc = qb_axpy_dev(n, alpha, dx, incx, (void *)dy, incy, 0);
  if (rc) return rc;
  cudaError_t e;
  if (sy) e = cudaMemcpy(y, dy, vec_bytes(n, incy), cudaMemcpyDeviceToHost); else e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(QB_ERR_CUDA, "qaxpy completion", e);
  return QB_OK;
}

/* ------------------------------------------------------------------ reference C ABI */
double quadblas_qdot(int n, void *x, int incx, void *y, int incy)
{
  clear_error();
  qb_quad r;
  if (qb_dot(n, x, incx, y, incy, &r)) return std::numeric_limits<double>::quiet_NaN();
  return qb_to_double(r); /* c_interface.hpp:30 */
}

double quadblas_qnrm2(int n, void *x, int incx)
{
  clear_error();
  qb_quad r;
  if (qb_nrm2(n, x, incx, &r)) return std::numeric_limits<double>::quiet_NaN();
  return qb_to_double(r); /* c_interface.hpp:42-43 */
}

void quadblas_qaxpy(int n, double alpha, void *x, int incx, void *y, int incy)
{
  clear_error();
  if (n <= 0) return; /* c_interface.hpp:49 */
  qb_quad a = qb_from_double(alpha);
  qb_axpy(n, &a, x, incx, y, incy);
}

void quadblas_qgemv(char layout, char trans, int m, int n, double alpha, void *A, int lda, void *x, int incx, double beta,
                    void *y, int incy)
{
  clear_error();
  qb_quad a = qb_from_double(alpha), b = qb_from_double(beta);
  bool col = is_col(layout);
  if (is_trans(trans)) { int t = m; m = n; n = t; col = !col; } /* c_interface.hpp:79-85 */
  qb_gemv(col ? 'C' : 'R', m, n, &a, A, lda, x, incx, &b, y, incy);
}

void quadblas_qgemm(char layout, char transa, char transb, int m, int n, int k, double alpha, void *A, int lda, void *B,
                    int ldb, double beta, void *C, int ldc)
{
  clear_error();
  qb_quad a = qb_from_double(alpha), b = qb_from_double(beta);
  qb_gemm(layout, transa, transb, m, n, k, &a, A, lda, B, ldb, &b, C, ldc);
}

} /* extern "C" */
