#!/usr/bin/env python
"""bench.py — headline benchmark of the binary128 hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--mode fast|ref] [--dist D113|D53]

Own arm (default): one "step" = one quadblas qgemm of the workload, device resident:
    N = 1 : C(SxS) = A(SxS) B(SxS), row-major, alpha=1, beta=0, S=8192 (BASELINE config 3)
    N > 1 : BASELINE config 4 sharding at fixed per-GPU work (weak scaling): C is (N*S x S), rank r owns
            the row block r; each step = NCCL broadcast of B (bytes) + local qgemm + gather of the C blocks, all
            inside the timed region (max over ranks).  --gather fused (default): the kernel that finishes the C
            elements stores them into every rank's copy over NVLink peer memory (no collective; a 4-byte all-reduce
            is the completion barrier); --gather nccl: NCCL all_gather (per row pass with --overlap P).
  --mode fast (default): QB_MODE_FAST -> the tensor-core path (csrc/qb_crt.cuh + qb_ozaki.cu): exact int8 residue
            planes, one tcgen05 kind::i8 GEMM per modulus, exact Chinese-remainder recombination, one rounding
            (--scheme digits: the digit-diagonal scheme).  BASELINE config 3's
            "integer-limb vs Ozaki" comparison: the integer-limb reference-order kernel is timed in
            extra.qgemm_reference_order (and is the headline with --mode ref).
  `value` = binary128 GFLOP/s (2mnk flops, benchmarks/benchmark.cpp:199-202) of the whole job.
  `e2e`   = same metric through the reference-named C entry point quadblas_qgemm with HOST (pinned)
            buffers: H2D of A, B, C and D2H of C inside the timed region.
  `roofline` = the dominant kernel: fast mode -> k_oz_mma against the tensor peak (2 x the measured bf16
            figure, int8 dense), its own duration from CUDA events the library records around each launch;
            ref mode -> k_gemm against the integer-issue ceiling measured live by the register-resident
            qFMA microbenchmark.  HBM-bound qgemv / qdot figures are in `extra`.
  `cpu_baseline` = the reference's own loops (oracle/_ref, libquadmath arithmetic) on the host cores,
            bounded sample.
Reference arm (--impl reference): times the reference's CPU implementation on the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


# ------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------ reference arm / cpu baseline
def time_reference_gemm(budget_s=12.0):
    """Reference loops (oracle/_ref: unmodified reference headers + libquadmath shim; falls back to the
    oracle port when the reference could not be compiled) on all host cores; bounded sample of the
    workload: qgemm m=512 (>= 500 so the reference goes parallel, level3.hpp:264) x s x s, row-major,
    alpha=1 beta=0, full-mantissa inputs.  Returns dict(value GFLOP/s, ...)."""
    import oracle_lib
    from qblas_b200 import quad
    ref = oracle_lib.load_ref()
    kind = "reference" if ref is not None else "port"
    orc = oracle_lib.load_oracle()
    cores = os.cpu_count() or 1
    if ref is not None:
        ref.set_num_threads(cores)
    rng = np.random.default_rng(42)
    m = 512

    def run(s):
        A = quad.random_quads(rng, m * s); B = quad.random_quads(rng, s * s); Cm = quad.random_quads(rng, m * s)
        t0 = time.perf_counter()
        if ref is not None:
            ref.gemm("R", m, s, s, 1.0, A, s, B, s, 0.0, Cm, s)
        else:
            orc.gemm("R", m, s, s, 1.0, A, s, B, s, 0.0, Cm, s)
        return time.perf_counter() - t0

    t = run(64)
    rate = 2.0 * m * 64 * 64 / t
    s = 64
    for cand in (96, 128, 192, 256, 384, 512, 768, 1024):
        if 2.0 * m * cand * cand / rate <= budget_s:
            s = cand
    t = run(s)
    return {"value": 2.0 * m * s * s / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind, "seconds": t,
            "sample": f"qgemm {m}x{s}x{s} row-major alpha=1 beta=0, D113 inputs, {('reference headers + libquadmath shim (SLEEF 3.8 unavailable offline)' if kind == 'reference' else 'oracle port')}, OMP threads={cores}, cpu={cpu_model()}"}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, args.warmup and 1)):
        time_reference_gemm(3.0)
    for _ in range(args.steps):
        vals.append(time_reference_gemm(8.0))
    best = sorted(vals, key=lambda d: d["value"])[len(vals) // 2]
    line = {
        "impl": "reference", "metric": "binary128 qgemm GFLOPS", "value": best["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "binary128 (software, libquadmath)", "data": "synthetic",
        "config": {"workload": "quadblas_qgemm row-major alpha=1 beta=0 (CPU: bounded sample of the 8192^3 workload)", "sample": best["sample"]},
        "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": best["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(json.dumps(line))


# ------------------------------------------------------------------ own arm
def _time_events(fn, reps):
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _int_issue_peak(qb, torch, dev):
    """register-resident qFMA microbenchmark (same qacc_fma as k_gemm, no global memory) -> binary128 GFLOP/s"""
    sink = torch.zeros((4, 2), dtype=torch.int64, device=dev)
    best = 0.0
    for variant, threads in ((104, 256), (102, 256), (104, 128), (4, 256), (2, 256)):
        qb.fma_microbench(variant, 148 * 4, threads, 64, sink)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nf = qb.fma_microbench(variant, 148 * 4, threads, 2000, sink); b.record(); torch.cuda.synchronize()
        best = max(best, nf / (a.elapsed_time(b) * 1e-3))
    return 2.0 * best / 1e9


def _secondary(qb, torch, dev, args, S, mode, extra):
    """BASELINE config 2 (qgemv / qdot vs HBM) and the reference-order integer-limb qgemm, reported in `extra`."""
    from gpu_util import dev_random
    peaks, src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    reps = 3
    try:
        if mode == qb.MODE_FAST:
            # reference-order (bit-exact) qgemm on the integer pipes, smaller cube: ~0.55 s per call at 4096^3
            Sr = 4096 if S >= 4096 else S
            Ar = dev_random((Sr * Sr,), args.dist, 21, dev); Br = dev_random((Sr * Sr,), args.dist, 22, dev); Cr = dev_random((Sr * Sr,), args.dist, 23, dev)
            qb.set_mode(qb.MODE_REFERENCE)
            qb.gemm("R", 256, 256, 256, 1.0, Ar, Sr, Br, Sr, 0.0, Cr, Sr)
            ms = _time_events(lambda: qb.gemm("R", Sr, Sr, Sr, 1.0, Ar, Sr, Br, Sr, 0.0, Cr, Sr), 1)
            pk = _int_issue_peak(qb, torch, dev)
            gf = 2.0 * Sr ** 3 / ms / 1e6
            extra["qgemm_reference_order"] = {
                "workload": f"quadblas_qgemm row-major {Sr}^3 alpha=1 beta=0, reference-order mode (bit exact vs the reference, kc=126), integer-limb kernel k_gemm",
                "ms": ms, "gflops": gf,
                "roofline": {"bound": "int-issue (IMAD/ALU pipes)", "achieved": gf, "peak": pk, "unit": "GFLOP/s (binary128)", "frac": gf / pk,
                             "peak_source": "live register-resident qFMA microbenchmark (same primitive, no memory)"}}
            qb.set_mode(mode)
            del Ar, Br, Cr
        mv = 32768 if S >= 8192 else 4096
        Av = dev_random((mv * mv,), args.dist, 11, dev); xv = dev_random((mv,), args.dist, 12, dev); yv = dev_random((mv,), args.dist, 13, dev)
        for md, name in ((qb.MODE_FAST, "fast"), (qb.MODE_REFERENCE, "reference")):
            qb.set_mode(md)
            qb.gemv("R", mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1)
            ms = _time_events(lambda: qb.gemv("R", mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1), reps)
            byt = 16.0 * (mv * mv + mv + 2 * mv)
            extra[f"qgemv_{name}"] = {"workload": f"quadblas_qgemv R/N {mv}x{mv} alpha=1 beta=0 ({name} mode)", "ms": ms, "gflops": 2.0 * mv * mv / ms / 1e6,
                                      "roofline": {"bound": "hbm", "achieved": byt / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": byt / ms / 1e6 / hbm, "peak_source": src}}
        del Av
        nd = 100_000_000 if S >= 8192 else 10_000_000
        xd = dev_random((nd,), args.dist, 14, dev); yd = dev_random((nd,), args.dist, 15, dev); res = torch.zeros((1, 2), dtype=torch.int64, device=dev)
        for md, name in ((qb.MODE_FAST, "fast"), (qb.MODE_REFERENCE, "reference")):
            qb.set_mode(md)
            if md == qb.MODE_REFERENCE:
                qb.quadblas_set_num_threads(4096)   # reference-order chunk count T (thread-count analogue)
            qb.dot(nd, xd, 1, yd, 1, res)
            ms = _time_events(lambda: qb.dot(nd, xd, 1, yd, 1, res), reps)
            extra[f"qdot_{name}"] = {"workload": f"qdot n={nd} unit stride ({name} mode" + (", T=4096)" if name == "reference" else ")"), "ms": ms,
                                     "roofline": {"bound": "hbm", "achieved": 32.0 * nd / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                                  "frac": 32.0 * nd / ms / 1e6 / hbm, "peak_source": src}}
        qb.quadblas_set_num_threads(0)
        del xd, yd
        # BASELINE config 1 (the reference README's benchmark, 0.06 GFLOPS there): quadblas_qgemm 1000^3, doubles cast to quad, alpha=1 beta=0,
        # through the reference-named C entry point with HOST buffers (synchronous, staging included) and device resident
        import time
        from gpu_util import to_host
        Sc = 1000
        A1 = dev_random((Sc * Sc,), "D53", 31, dev); B1 = dev_random((Sc * Sc,), "D53", 32, dev); C1 = dev_random((Sc * Sc,), "D53", 33, dev)
        hA1, hB1, hC1 = to_host(A1), to_host(B1), to_host(C1)
        for md, name in ((qb.MODE_FAST, "fast"), (qb.MODE_REFERENCE, "reference")):
            qb.set_mode(md)
            qb.gemm("R", Sc, Sc, Sc, 1.0, A1, Sc, B1, Sc, 0.0, C1, Sc)
            ms = _time_events(lambda: qb.gemm("R", Sc, Sc, Sc, 1.0, A1, Sc, B1, Sc, 0.0, C1, Sc), reps)
            qb.quadblas_qgemm("R", "N", "N", Sc, Sc, Sc, 1.0, hA1, Sc, hB1, Sc, 0.0, hC1, Sc)
            t0 = time.perf_counter()
            for _ in range(reps):
                qb.quadblas_qgemm("R", "N", "N", Sc, Sc, Sc, 1.0, hA1, Sc, hB1, Sc, 0.0, hC1, Sc)
            ms_host = (time.perf_counter() - t0) * 1e3 / reps
            extra[f"qgemm_cfg1_1000_{name}"] = {
                "workload": f"quadblas_qgemm('R','N','N',1000,1000,1000, 1.0, A,1000, B,1000, 0.0, C,1000), doubles U(-1,1) cast to quad ({name} mode)",
                "device_resident": {"ms": ms, "gflops": 2.0 * Sc ** 3 / ms / 1e6},
                "host_buffers_c_abi": {"ms": ms_host, "gflops": 2.0 * Sc ** 3 / ms_host / 1e6, "note": "pageable numpy buffers, staging copies included, wall clock"},
                "plan": qb.oz_last_stats() if md == qb.MODE_FAST else None,
                "reference_readme_gflops": 0.06}
        del A1, B1, C1
    except Exception as e:  # secondary figures must never take the headline down
        extra["error"] = repr(e)
    qb.set_mode(mode)


def own_arm(args, rank, world, local_rank):
    import torch
    import qblas_b200 as qb
    from gpu_util import dev_random, to_host
    from qblas_b200 import quad
    import oracle_lib

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    qb.init()
    mode = qb.MODE_FAST if args.mode == "fast" else qb.MODE_REFERENCE
    qb.set_mode(mode)
    if args.keep is not None:
        qb.set_tensor_keep(args.keep)
    if args.scheme is not None:
        qb.set_tensor_scheme(1 if args.scheme == "residues" else 0)
    if args.pass_shape:
        qb.set_tensor_pass_shape(args.pass_shape)
    if args.host_slabs:
        qb.set_host_slabs(args.host_slabs)
    S = args.size
    m_loc, n, k = S, S, S
    M = S * world
    strong = False
    if args.shape:   # BASELINE config 4: a FIXED global problem, C row-blocks of M / world rows per GPU (strong scaling)
        M, n, k = (int(v) for v in args.shape.split(","))
        assert M % world == 0, "--shape M must be a multiple of the number of GPUs"
        m_loc = M // world
        strong = True
    dev = torch.device("cuda", local_rank)

    # inputs resident in HBM before the timed region (3 x 1 GiB at S=8192: far larger than the 126 MB L2)
    A = dev_random((m_loc * k,), args.dist, 100 + rank, dev)
    B = dev_random((k * n,), args.dist, 7, dev) if rank == 0 or world == 1 else torch.empty((k * n, 2), dtype=torch.int64, device=dev)
    Cfull = dev_random((M * n,), args.dist, 9, dev)          # C_in (beta=0 still reads it, level3.hpp:107)
    peerbuf = None
    if world > 1 and args.gather == "fused" and mode == qb.MODE_FAST:
        # fused gather: C_full lives in peer-mapped memory; the reconstruction kernel stores every finished element into all
        # ranks' copies over NVLink (qb_set_gemm_peer_outputs), so no all-gather re-reads and re-sends the blocks
        from qblas_b200 import dist as qd
        try:
            peerbuf = qd.PeerBuffer(M * n * 16)
            Cp = peerbuf.tensor.view(torch.int64).reshape(M * n, 2)
            Cp.copy_(Cfull)
            Cfull = Cp
            torch.cuda.empty_cache()     # hand the first copy back to the driver: the library sizes its workspace from cudaMemGetInfo
        except Exception as e:   # no peer access on this box: NCCL gather
            print(f"[bench] fused gather unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr, flush=True)
            peerbuf = None
    Cblk = Cfull[rank * m_loc * n:(rank + 1) * m_loc * n]

    Cb3 = Cblk.reshape(m_loc, n, 2)
    Cf3 = Cfull.reshape(world, m_loc, n, 2)

    def gemm_and_gather():
        """local qgemm + all-gather of the C blocks.  With --overlap P > 1 the rows are produced in P passes and the gather of
        each pass's rows is issued from the library's row-pass hook (qb_set_gemm_pass_callback), so it runs on NCCL's stream
        while the next pass computes; only the last pass's gather is exposed."""
        if world == 1:
            qb.gemm("R", m_loc, n, k, 1.0, A, k, B, n, 0.0, Cblk, n)
            return
        if fused["on"]:
            qb.gemm("R", m_loc, n, k, 1.0, A, k, B, n, 0.0, Cblk, n)   # peer outputs are set: the fold kernel writes all ranks' copies
            dist.all_reduce(fused["token"])                            # completion barrier of the peer stores, in stream order
            return
        if args.overlap <= 1:
            qb.gemm("R", m_loc, n, k, 1.0, A, k, B, n, 0.0, Cblk, n)
            dist.all_gather_into_tensor(Cfull, Cblk)          # in place: Cblk is rank's slice of Cfull
            return
        works = []

        def on_rows(r0, rows):
            works.append(dist.all_gather([Cf3[q, r0:r0 + rows] for q in range(world)], Cb3[r0:r0 + rows], async_op=True))

        qb.set_gemm_pass_callback(on_rows, args.overlap)
        try:
            qb.gemm("R", m_loc, n, k, 1.0, A, k, B, n, 0.0, Cblk, n)
        finally:
            qb.set_gemm_pass_callback(None)
        for w in works:
            w.wait()

    def step():
        if world > 1:
            dist.broadcast(B, src=0)                          # byte-typed payload (int64 view of quads)
        gemm_and_gather()

    fused = {"on": False, "token": torch.zeros(1, dtype=torch.int32, device=dev)}
    if peerbuf is not None:
        # one probing step: every rank must have run the path with the fused stores, otherwise all of them use the NCCL gather
        qb.set_gemm_peer_outputs([peerbuf.ptrs[q] + rank * m_loc * n * 16 for q in range(world) if q != rank])
        fused["on"] = True
        step()
        ok = torch.tensor([1 if qb.gemm_peer_written() == world - 1 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fused["on"] = False
            qb.set_gemm_peer_outputs(None)
    args.gather_used = "fused" if fused["on"] else "nccl"

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = qb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    mma_ms, mma_launches = [], 0
    ev0.record()
    for i in range(args.steps):
        if world > 1:
            dist.broadcast(B, src=0)
        kev[i][0].record()
        gemm_and_gather()
        kev[i][1].record()
        if mode == qb.MODE_FAST and rank == 0 and i == args.steps - 1:
            pass  # per-kernel events are read after the timed region (reading them blocks the host)
    ev1.record()
    sync()
    if fused["on"]:
        qb.set_gemm_peer_outputs(None)    # nothing after the timed region may write into the peers
    # every rank must now hold every rank's block: compare per-block checksums of the local C_full with the owners' own
    args.gather_bad = 0
    if world > 1:
        sums = Cfull.view(torch.int64).reshape(world, -1).sum(dim=1)            # wrapping int64 sums, one per block
        owners = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(owners, sums[rank:rank + 1].clone())
        args.gather_bad = int((sums != torch.cat(owners)).sum().item())
    launches = qb.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    call_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    flops_step = 2.0 * M * n * k
    value = flops_step / (ms_step * 1e-3) / 1e9
    plan = qb.oz_last_stats() if mode == qb.MODE_FAST else None
    if mode == qb.MODE_FAST:
        # the library brackets every tcgen05 launch of the LAST timed qgemm with CUDA events on its stream
        ms_k, nl = qb.oz_last_mma_ms()
        mma_ms, mma_launches = ms_k, nl

    # ---- parity of the timed result, sampled (outside the timed region)
    orc = oracle_lib.load_oracle()
    rng = np.random.default_rng(5 + rank)
    ns = 48 if mode == qb.MODE_REFERENCE else 16
    idx = np.stack([rng.integers(0, m_loc, ns), rng.integers(0, n, ns)], axis=1)
    ri, ci = torch.as_tensor(idx[:, 0], device=dev), torch.as_tensor(idx[:, 1], device=dev)
    got = to_host(Cblk.reshape(m_loc, n, 2)[ri, ci].contiguous())
    # only the sampled rows of A and columns of B leave the device: a (ns x k) by (k x ns) problem whose
    # diagonal holds the sampled entries (same k order, so the reference-order oracle applies unchanged)
    Ah = to_host(A.reshape(m_loc, k, 2)[ri].contiguous().reshape(ns * k, 2))
    Bh = to_host(B.reshape(k, n, 2)[:, ci].contiguous().reshape(k * ns, 2))
    didx = np.stack([np.arange(ns), np.arange(ns)], axis=1)
    if mode == qb.MODE_REFERENCE:
        # C_in = None -> +0 in the oracle; mul(0, c_in) = +-0 and fma(alpha, s, +-0) == s unless s == 0
        exp = orc.gemm_sample("R", ns, ns, k, 1.0, Ah, k, Bh, ns, 0.0, None, ns, didx)
        mism = int((~quad.same_bits(got, exp)).sum())
        against = "oracle/qoracle.c, reference order, bit exact"
    else:
        from fractions import Fraction
        from exact_ref import exact_matmul_rounded
        mism = 0
        exp_ref = orc.gemm_sample("R", ns, ns, k, 1.0, Ah, k, Bh, ns, 0.0, None, ns, didx)
        ab = orc.absdot_sample("R", k, Ah, k, Bh, ns, didx)
        u = Fraction(1, 2 ** 113); gam = k * u / (1 - k * u)
        f = lambda v: quad.to_fraction(int(v[1]), int(v[0]))
        for q in range(ns):
            if _plan_exact(plan):   # residue scheme / all diagonals: the exact inner product rounded once
                s_ex = exact_matmul_rounded(Ah[q * k:(q + 1) * k], k, np.ascontiguousarray(Bh[q::ns][:k]), 1, 1, 1, k)
                if not quad.same_bits(got[q:q + 1], s_ex).all():
                    mism += 1
            if abs(f(got[q]) - f(exp_ref[q])) > 2 * gam * f(ab[q]):   # and always inside the fast-mode contract vs the reference order
                mism += 1
        against = ("exact big-integer inner product rounded once (bit exact) AND " if _plan_exact(plan) else "") + \
            "gamma_k(|A||B|) bound vs the reference-order oracle"
    del Ah, Bh

    extra = {}
    roof = None
    e2e = None
    cpu = None
    if rank == 0:
        peaks, src = measured_peaks()
        if mode == qb.MODE_FAST and plan and plan["pairs"] > 0:
            int8_ops = plan["pairs"] * 2.0 * m_loc * n * plan["Kp"]
            tops = int8_ops / (mma_ms * 1e-3) / 1e12
            bf16_sus = float(peaks.get("bf16_tflops_sustained", 1400.0)); bf16_burst = float(peaks.get("bf16_tflops", 1590.0))
            # residue scheme: a few ms per launch at (nearly) full clocks -> the burst figure; digit diagonals: 30-60 ms launches under the power cap -> sustained
            peak = 2.0 * (bf16_burst if plan.get("scheme") == "residues" else bf16_sus)
            # DRAM traffic per launch of k_oz_mma from the committed ncu --set full capture of THIS configuration
            # (profiles/r1g_oz_mma_8192_ncu_full.txt: dram__bytes_read.sum 38.95 GB + dram__bytes_write.sum 4.25 GB); null for any other plan
            residues = plan.get("scheme") == "residues"
            profiled = (not residues) and (m_loc, n, plan["Kp"], plan["SA"], plan["SB"], plan["keep"], plan["nchunks"]) == (8192, 8192, 8192, 18, 18, 16, 2)
            if residues:
                alg = (f"one binary128 flop = {plan['pairs']} int8 ops: row/column block fixed point (W_A = {plan['WA']}, W_B = {plan['WB']} bits), "
                       f"one int8 GEMM per modulus, {plan['pairs']} pairwise coprime moduli <= 256 (product > 2 k 2^(W_A+W_B)), exact CRT reconstruction")
            else:
                alg = (f"one binary128 flop = {plan['pairs']} int8 ops ({plan['SA']}x{plan['SB']} signed-digit slices, {plan['keep']} of {plan['ndiag']} "
                       "diagonals multiplied)")
            roof = {"bound": "tensor", "kernel": "k_oz_mma (tcgen05.mma kind::i8, TMA-fed, TMEM accumulators)", "achieved": tops, "peak": peak,
                    "unit": "TFLOP/s", "frac": tops / peak,
                    "traffic": 43.2e9 if profiled else (4.02e9 * plan["pairs"] / 40.0 if residues and (m_loc, n, plan["Kp"], plan["row_passes"]) == (8192, 8192, 8192, 4) else None),
                    "traffic_note": ("bytes per launch (4 launches per qgemm, one per pass of 2048 rows), from profiles/r1h_crt_mma_8192_ncu_full.txt (40 moduli: dram read 3.36 GB + write 0.66 GB, "
                                     "scaled to this plan's moduli): equal to the algorithmic bytes N*(2048*Kp + n*Kp) of int8 planes read once + N*2048*n residue bytes written "
                                     "(1.1 TB/s = 17% of HBM while the tensor pipe is 88% active)") if residues else "bytes per launch (2 launches per qgemm); algorithmic operand + result bytes per launch = 1.2 GB of digit planes + 4.3 GB of int32 diagonals: the planes are re-read once per digit-plane pair through L2 (1.5 TB/s = 23% of HBM while the tensor pipe is 85% active - not the bound)" if profiled else None,
                    "peak_source": (f"2 x MEASURED_PEAKS.json bf16_tflops (burst {bf16_burst}; sustained {bf16_sus}) [{src}]: int8 dense issues at twice the bf16 rate on sm_100a, "
                                    "no int8 figure is driver-measured; burst because each launch lasts a few ms (measured 3177 TOPS when timed alone)") if plan.get("scheme") == "residues" else
                                   (f"2 x MEASURED_PEAKS.json bf16_tflops_sustained ({bf16_sus}; burst {bf16_burst}) [{src}]: int8 dense issues at twice the bf16 rate on sm_100a, "
                                    "no int8 figure is driver-measured; sustained because the kernel runs inside a long back-to-back step"),
                    "algorithmic": f"{alg}: "
                                   f"{plan['pairs']} x 2*m*n*Kp = {int8_ops:.4g} int8 ops per qgemm in {mma_launches} launch(es) of k_oz_mma, "
                                   f"{mma_ms:.2f} ms summed (CUDA events on the launching stream, last timed step); whole qgemm call {call_ms:.2f} ms",
                    "kernel_ms": mma_ms, "kernel_share_of_step": mma_ms / call_ms, "plan": plan,
                    "binary128_gflops_in_kernel": 2.0 * m_loc * n * k / (mma_ms * 1e-3) / 1e9}
        else:
            pk = _int_issue_peak(qb, torch, dev)
            kern_gflops = 2.0 * m_loc * n * k / (call_ms * 1e-3) / 1e9
            roof = {"bound": "int-issue (IMAD/ALU pipes; not hbm, not tensor)", "kernel": "k_gemm", "achieved": kern_gflops, "peak": pk,
                    "unit": "GFLOP/s (binary128)", "frac": kern_gflops / pk, "traffic": None,
                    "peak_source": "live register-resident qFMA microbenchmark (qb_fma_microbench_dev, best of 5 shapes; same qacc_fma as k_gemm, no global memory)",
                    "algorithmic": f"2*m*n*k = {2.0 * m_loc * n * k:.4g} binary128 flops per launch; avg launch {call_ms:.2f} ms (CUDA events)"}
        if not args.no_extra and not strong:
            _secondary(qb, torch, dev, args, S, mode, extra)

    # ---- e2e: reference-named C entry point with HOST buffers (pinned), copies inside the timed region
    if strong:   # config-4 tool runs report the device-resident + collective figure only (B alone is 16 GiB of pinned host memory)
        _print_line(args, rank, world, dist, torch, dev, qb, mode, plan, value, ms_step, M, n, k, m_loc, None, launches, clk, roof, None, ns, mism, against,
                    call_ms, extra, strong)
        return
    torch.cuda.empty_cache()
    hA = torch.empty((m_loc * k, 2), dtype=torch.int64).pin_memory(); hA.copy_(A)
    hB = torch.empty((k * n, 2), dtype=torch.int64).pin_memory(); hB.copy_(B)
    hC = torch.empty((m_loc * n, 2), dtype=torch.int64).pin_memory(); hC.copy_(Cblk)
    nA, nB, nC = (x.numpy().view(np.uint64) for x in (hA, hB, hC))
    e2e_steps = max(1, min(args.steps, 3))
    qb.quadblas_qgemm("R", "N", "N", m_loc, n, k, 1.0, nA, k, nB, n, 0.0, nC, n)  # warm (allocates staging)
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        qb.quadblas_qgemm("R", "N", "N", m_loc, n, k, 1.0, nA, k, nB, n, 0.0, nC, n)  # synchronous: returns with C on the host
    t1 = time.perf_counter()
    te = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = 2.0 * M * n * k / float(te.item()) / 1e9
    e2e = {"value": e2e_val, "unit": "GFLOP/s", "h2d_bytes_per_step": 16 * (m_loc * k + k * n + m_loc * n), "d2h_bytes_per_step": 16 * m_loc * n,
           "ms_per_step": float(te.item()) * 1e3,
           "api": "quadblas_qgemm (reference C ABI), pinned host buffers, synchronous; per-rank row block, no collective", "steps": e2e_steps}
    _print_line(args, rank, world, dist, torch, dev, qb, mode, plan, value, ms_step, M, n, k, m_loc, e2e, launches, clk, roof, "time", ns, mism, against,
                call_ms, extra, strong)


def _plan_exact(plan):
    return bool(plan) and plan["pairs"] > 0 and (plan.get("scheme") == "residues" or plan["keep"] >= plan["ndiag"])


def _mode_text(plan):
    if plan and plan.get("scheme") == "residues":
        return (f"fast: exact int8 residue planes on tcgen05 ({plan['pairs']} moduli, one GEMM each) + Chinese-remainder reconstruction: "
                "inner products exact, rounded once")
    return "fast: Ozaki-style exact int8 slicing on tcgen05; " + (
        f"{plan['keep']} leading diagonals + per-element check/fix-up ({plan['flagged']} entries fixed, {plan['redo_passes']} passes redone): inside the gamma_k bound"
        if plan and plan["keep"] < plan["ndiag"] else "all diagonals: inner products exact, rounded once")


def _print_line(args, rank, world, dist, torch, dev, qb, mode, plan, value, ms_step, M, n, k, m_loc, e2e, launches, clk, roof, cpu, ns, mism, against,
                call_ms, extra, strong):
    mism_t = torch.tensor([mism, getattr(args, "gather_bad", 0)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(mism_t)

    if rank == 0:
        if cpu == "time":
            try:
                cpu = time_reference_gemm(12.0)
                cpu = {k2: cpu[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        fast = mode == qb.MODE_FAST
        line = {
            "metric": "binary128 qgemm GFLOPS", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": ("binary128 (exact int8 residues on the tensor cores, Chinese-remainder recombination, one rounding)" if plan and plan.get("scheme") == "residues"
                      else "binary128 (exact signed 8-bit slices on the int8 tensor cores, wide-integer recombination, one rounding)") if fast
                     else "binary128 (software, u32 integer limbs)",
            "data": "synthetic",
            "config": {"workload": f"quadblas_qgemm row-major {M}x{n}x{k} alpha=1 beta=0 ({'C row-blocks of ' + str(m_loc) + ' rows per GPU, ' + ('NCCL broadcast(B) + fused gather of C (the reconstruction kernel stores each finished element into every rank over NVLink peer memory; barrier) in the timed region' if getattr(args, 'gather_used', 'nccl') == 'fused' else 'NCCL broadcast(B)+all_gather(C) in the timed region' + (f', all-gather issued per row pass ({args.overlap} passes) from the row-pass hook' if args.overlap > 1 else '')) if world > 1 else 'BASELINE config 3, 1xB200' if not strong else 'BASELINE config 4 shape on 1 GPU'})",
                       "mode": _mode_text(plan) if fast
                               else "reference-order (bit exact, kc=126), integer-limb kernel",
                       "inputs": f"{args.dist}: full 113-bit random mantissas, device resident" if args.dist != "D53" else "D53: doubles U(-1,1) cast to quad (the reference's own benchmark distribution)",
                       "l2": "inputs (3 x 1 GiB at 8192^3) and the int8 planes (GBs) exceed the 126 MB L2; no flush needed", "parallelism": f"row-block x{world}"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "parity": {"checked_entries": ns * world, "mismatches": int(mism_t[0].item()), "against": against,
                       "gathered_blocks_checked": world * world if world > 1 else 0, "gathered_blocks_wrong": int(mism_t[1].item())},
            "call_ms": call_ms, "extra": extra,
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE line (the JSON): everything any library prints to fd 1 (NCCL's version banner, ...) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--shape", default=None, help="M,N,K of a fixed global problem split into M/world-row blocks (BASELINE config 4, e.g. 32768,32768,32768)")
    ap.add_argument("--mode", default="fast", choices=["ref", "fast"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary qgemv/qdot/reference-order figures")
    ap.add_argument("--dist", default="D113", choices=["D113", "D53", "Dexp"])
    ap.add_argument("--overlap", type=int, default=4, help="N > 1: row passes whose all-gathers overlap the next pass (1 = one all-gather after the qgemm)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N > 1: C blocks reach the other ranks by peer stores from the kernel that finishes them (fused) or by NCCL all-gather")
    ap.add_argument("--host-slabs", type=int, default=0, help="e2e: C slabs of the pipelined all-host qgemm (library default 4)")
    ap.add_argument("--pass-shape", type=int, default=0, choices=[0, 1], help="residue scheme row passes: 0 equal (default), 1 short first / last pass (experimental)")
    ap.add_argument("--scheme", default=None, choices=["residues", "digits"], help="tensor path: residue planes + CRT (library default) or digit diagonals")
    ap.add_argument("--keep", type=int, default=None, help="tensor path: leading diagonals multiplied (0 = all = exact inner products; default: library default 16)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        _claim_stdout()
        reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run as the contract describes
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    _claim_stdout()
    own_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
