#!/usr/bin/env python
"""bench.py — headline benchmark of the binary128 hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S] [--mode fast|ref] [--dist D113|D53|Dexp]

Own arm (default): one "step" = one quadblas qgemm of the workload, device resident:
    N = 1 : C(SxS) = A(SxS) B(SxS), row-major, alpha=1, beta=0, S=8192 (BASELINE config 3)
    N > 1 : BASELINE config 4 sharding at fixed per-GPU work (weak scaling): C is (N*S x S), rank r owns the row block r; each
            step = qblas_b200.dist.qgemm_row_sharded: B travels from rank 0 to every rank and the C blocks reach every rank, all
            inside the timed region (max over ranks).  --bcast panels (default): B is broadcast in packed column panels DURING
            the product (qb_set_gemm_b_panels); --bcast whole: one NCCL broadcast before it.  --gather fused (default): the kernel
            that finishes the C elements stores them into every rank's copy over NVLink (one store to the NVSwitch multicast
            address of torch symmetric memory when available, else one store per peer; a 4-byte all-reduce is the completion
            barrier); --gather nccl: NCCL all_gather (per row pass with --overlap P).
    extra.cfg4_strong (every N): BASELINE config 4 itself, 32768^3 as a FIXED global problem cut into C row-blocks of 32768 / N
            rows, same step; the N = 1 line carries the single-GPU time, so the per-N lines give the strong-scaling curve.
  --mode fast (default): QB_MODE_FAST -> the tensor-core path (csrc/qb_crt.cuh + qb_ozaki.cu): exact int8 residue planes, one
            tcgen05 kind::i8 GEMM per modulus, exact Chinese-remainder recombination, one rounding.  BASELINE config 3's
            "integer-limb vs Ozaki" comparison: extra.cfg3_* time both kernels on D53 / D113 / Dexp inputs at 8192^3, and the
            integer-limb reference-order kernel is timed in extra.qgemm_reference_order (the headline with --mode ref).
  `value` = binary128 GFLOP/s (2mnk flops, benchmarks/benchmark.cpp:199-202) of the whole job.
  `e2e`   = same metric through the reference-named C entry point quadblas_qgemm with HOST (pinned)
            buffers: H2D of A, B, C and D2H of C inside the timed region.
  `roofline` = the dominant kernel: fast mode -> k_oz_mma against the tensor peak (2 x the measured bf16
            figure, int8 dense), its own duration from CUDA events the library records around each launch;
            ref mode -> k_gemm_nb against the dispatch clocks of its exact 113x113-bit product (16 IMAD.WIDE); the register-resident
            qFMA microbenchmark of the first-version step is reported next to it.  HBM-bound qgemv / qdot figures are in `extra`.
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
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


def _int8_peak(qb, torch, dev):
    """The tcgen05 kind::i8 pipeline of k_oz_mma on operands that stay in L2 (2 x 32 MB of int8 planes, 64 plane pairs per launch,
    ~0.4 ms per launch): the measured int8 rate of this kernel's own instruction stream with HBM reads out of the way -> TOPS."""
    S, m, n, Kp = 8, 2048, 2048, 2048
    g = torch.Generator(device=dev); g.manual_seed(1)
    pa = torch.randint(-128, 128, (S, m, Kp), generator=g, device=dev, dtype=torch.int8)
    pb = torch.randint(-128, 128, (S, n, Kp), generator=g, device=dev, dtype=torch.int8)
    D = torch.zeros((2 * S - 1, m, n), device=dev, dtype=torch.int32)
    for _ in range(3):
        qb.oz_i8gemm(pa, pb, m, n, D)
    ms = _time_events(lambda: qb.oz_i8gemm(pa, pb, m, n, D), 20)
    return S * S * 2.0 * m * n * Kp / (ms * 1e-3) / 1e12, ms


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
            qb.set_ref_gemm_kernel(0)                      # the first version of the kernel, same bits (side by side)
            try:
                qb.gemm("R", 256, 256, 256, 1.0, Ar, Sr, Br, Sr, 0.0, Cr, Sr)
                ms_old = _time_events(lambda: qb.gemm("R", Sr, Sr, Sr, 1.0, Ar, Sr, Br, Sr, 0.0, Cr, Sr), 1)
            finally:
                qb.set_ref_gemm_kernel(1)
            pk = _int_issue_peak(qb, torch, dev)
            gf = 2.0 * Sr ** 3 / ms / 1e6
            sm_ghz = 1.965
            clk = ms * 1e-3 * sm_ghz * 1e9 * 148 * 4 / (Sr ** 3 / 32.0)     # dispatch clocks per warp-qFMA and SM sub-partition at the boost clock
            floor_gf = 2.0 * 32.0 * 148 * 4 * sm_ghz / 69.0                   # GFLOP/s if the step were nothing but its 16 wide multiplies
            extra["qgemm_reference_order"] = {
                "workload": f"quadblas_qgemm row-major {Sr}^3 alpha=1 beta=0, reference-order mode (bit exact vs the reference, kc=126), integer-limb kernel k_gemm_nb (branch-free step, staged decoded operands)",
                "ms": ms, "gflops": gf,
                "first_version_k_gemm": {"ms": ms_old, "gflops": 2.0 * Sr ** 3 / ms_old / 1e6},
                "clk_per_warp_qfma_per_subpartition": clk,
                "clk_model": "sum of dispatch clocks over the SASS of the hot block (tools/sass_cost.py, calibrated by tools/exp/mb_pipes.cu): 252 for k_gemm_nb, 273 for k_gemm; the 16 IMAD.WIDE of the 113x113-bit product alone are 69",
                "roofline": {"bound": "integer dispatch (the sum of the dispatch clocks of the step's instruction stream; no single pipe and no memory level is saturated)",
                             "achieved": gf, "peak": floor_gf, "unit": "GFLOP/s (binary128)", "frac": gf / floor_gf,
                             "peak_source": "the exact 113x113-bit product alone: 16 IMAD.WIDE x 4.3 dispatch clocks = 69 clocks per warp-qFMA and SM sub-partition at the boost clock; alignment, signed add, normalisation and rounding are what the kernel adds on top",
                             "register_resident_microbench_first_version_gflops": pk, "frac_of_that_microbench": gf / pk,
                             "ncu": "profiles/r2_kgemm_nb_ncu_full.txt: ALU pipe 73 %, fmaheavy 46 %, issue slots 62 %, DRAM < 1 %"}}
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
            if md == qb.MODE_FAST:
                # the sliced FP64 kernel (default) declines rows it cannot guarantee; col-major takes the same kernel; variant 1 = the
                # window accumulator of round 1 on both layouts, for comparison
                extra["qgemv_fast"]["kernel"] = "k_gemv_f64 (sliced FP64 accumulate, csrc/qslice.cuh)"
                extra["qgemv_fast"]["rows_declined"] = qb.gemv_last_declined()
                qb.gemv("C", mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1)
                msc = _time_events(lambda: qb.gemv("C", mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1), reps)
                extra["qgemv_fast_colmajor"] = {"workload": f"quadblas_qgemv C/N {mv}x{mv} alpha=1 beta=0 (fast mode)", "ms": msc,
                                                "roofline": {"bound": "hbm", "achieved": byt / msc / 1e6, "peak": hbm, "unit": "GB/s", "frac": byt / msc / 1e6 / hbm, "peak_source": src}}
                qb.set_fast_variant(1)
                try:
                    w = {}
                    for lay in "RC":
                        qb.gemv(lay, mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1)
                        w[lay] = _time_events(lambda: qb.gemv(lay, mv, mv, 1.0, Av, mv, xv, 1, 0.0, yv, 1), reps)
                    extra["qgemv_fast_window_variant"] = {"what": "qb_set_fast_variant(1): the 192-bit window accumulate (integer multiplier) on the same call",
                                                          "ms_row_major": w["R"], "frac_row_major": byt / w["R"] / 1e6 / hbm,
                                                          "ms_col_major": w["C"], "frac_col_major": byt / w["C"] / 1e6 / hbm}
                finally:
                    qb.set_fast_variant(2)
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
            if md == qb.MODE_FAST:
                extra["qdot_fast"]["kernel"] = "k_dot_wide_tma (192-bit window accumulate, tiles of 512 pairs by cp.async.bulk in a two-stage ring; one launch incl. the fold)"
                for n2 in (10_000_000,):
                    ms2 = _time_events(lambda: qb.dot(n2, xd, 1, yd, 1, res), 20)   # back to back: the launch latency of the first call is amortised
                    extra[f"qdot_fast_n{n2}"] = {"ms": ms2, "gbs": 32.0 * n2 / ms2 / 1e6, "frac": 32.0 * n2 / ms2 / 1e6 / hbm}
                qb.nrm2(nd, xd, 1, res)
                ms3 = _time_events(lambda: qb.nrm2(nd, xd, 1, res), reps)
                extra["qnrm2_fast"] = {"workload": f"qnrm2 n={nd} unit stride (fast mode)", "ms": ms3, "kernel": "k_sumsq_tma (sliced FP64 sum of squares fed by cp.async.bulk, csrc/qslice.cuh)",
                                       "roofline": {"bound": "hbm", "achieved": 16.0 * nd / ms3 / 1e6, "peak": hbm, "unit": "GB/s", "frac": 16.0 * nd / ms3 / 1e6 / hbm, "peak_source": src}}
                qb.set_fast_variant(1)
                try:
                    qb.nrm2(nd, xd, 1, res)
                    ms4 = _time_events(lambda: qb.nrm2(nd, xd, 1, res), reps)
                    extra["qnrm2_fast"]["window_variant_ms"] = ms4
                    extra["qnrm2_fast"]["window_variant_frac"] = 16.0 * nd / ms4 / 1e6 / hbm
                finally:
                    qb.set_fast_variant(2)
        qb.quadblas_set_num_threads(0)
        # qaxpy (order-free: one kernel for both modes, every y_i = one correctly rounded FMA): 48 bytes per element
        qb.axpy(nd, 1.5, xd, 1, yd, 1)
        msa = _time_events(lambda: qb.axpy(nd, 1.5, xd, 1, yd, 1), reps)
        extra["qaxpy"] = {"workload": f"qaxpy n={nd} unit stride (bit exact in both modes)", "ms": msa, "gqfma_per_s": nd / msa / 1e6,
                          "roofline": {"bound": "hbm", "achieved": 48.0 * nd / msa / 1e6, "peak": hbm, "unit": "GB/s", "frac": 48.0 * nd / msa / 1e6 / hbm, "peak_source": src}}
        del xd, yd
        # BASELINE config 1 (the reference README's benchmark, 0.06 GFLOPS there): quadblas_qgemm 1000^3, doubles cast to quad, alpha=1 beta=0,
        # through the reference-named C entry point with HOST buffers (synchronous, staging included) and device resident
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
        # the headline workload through the reference C ABI from PAGEABLE host arrays (what std::vector<Sleef_quad> / numpy callers of the
        # reference pass; `e2e` itself uses page-locked buffers): feeder threads + page-locked rings inside the library
        if S >= 8192:
            qb.set_mode(qb.MODE_FAST)
            hA = to_host(dev_random((S * S,), args.dist, 41, dev)); hB = to_host(dev_random((S * S,), args.dist, 42, dev)); hC = to_host(dev_random((S * S,), args.dist, 43, dev))
            qb.quadblas_qgemm("R", "N", "N", S, S, S, 1.0, hA, S, hB, S, 0.0, hC, S)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter(); qb.quadblas_qgemm("R", "N", "N", S, S, S, 1.0, hA, S, hB, S, 0.0, hC, S); ts.append(time.perf_counter() - t0)
            extra["e2e_pageable_host_buffers"] = {"workload": f"quadblas_qgemm row-major {S}^3 alpha=1 beta=0 (fast mode), pageable numpy arrays, synchronous, wall clock",
                                                  "ms": min(ts) * 1e3, "gflops": 2.0 * S ** 3 / min(ts) / 1e9,
                                                  "note": "a plain cudaMemcpyAsync pipeline from pageable memory took 275 ms for this call"}
            del hA, hB, hC
    except Exception as e:  # secondary figures must never take the headline down
        extra["error"] = repr(e)
    qb.set_mode(mode)


def _cfg3(qb, torch, dev, S, extra):
    """BASELINE config 3 / SURVEY §8d cfg3: the three value sets on BOTH fast-mode kernels at S^3 — the tensor path and the integer-limb
    kernel (qb_set_tensor_path(0)) — with which path ran and a sampled contract check of each."""
    import oracle_lib
    from gpu_util import dev_random, to_host
    orc = oracle_lib.load_oracle()
    qb.set_mode(qb.MODE_FAST)
    try:
        for kind in ("D53", "D113", "Dexp"):
            A = dev_random((S * S,), kind, 41, dev); B = dev_random((S * S,), kind, 42, dev); C = dev_random((S * S,), kind, 43, dev)
            qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
            ms = _time_events(lambda: qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S), 3)
            plan = qb.oz_last_stats()
            rng = np.random.default_rng(9)
            ri = np.sort(rng.choice(S, 32, replace=False)); ci = np.sort(rng.choice(S, 32, replace=False))
            par = _parity_grid(torch, orc, A, S, B, S, C, S, S, ri, ci, exact_bits=plan["exact"])
            entry = {"workload": f"quadblas_qgemm row-major {S}^3 alpha=1 beta=0, {kind} inputs, fast mode", "tensor_path": {
                "ran": "k_oz_mma (tcgen05 residue scheme)" if plan["pairs"] > 0 else "declined -> integer-limb kernel", "ms": ms, "gflops": 2.0 * S ** 3 / ms / 1e6,
                "plan": plan, "parity": par}}
            if kind == "Dexp":   # the integer-limb kernel on the same buffers (data independent to first order: timed once, on the set that used to fall back to it)
                qb.set_tensor_path(qb.TENSOR_OFF)
                ms_i = _time_events(lambda: qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S), 1)
                qb.set_tensor_path(qb.TENSOR_AUTO)
                entry["integer_limb_kernel"] = {"ran": "k_gemm (single chain per element)", "ms": ms_i, "gflops": 2.0 * S ** 3 / ms_i / 1e6,
                                                "parity": _parity_grid(torch, orc, A, S, B, S, C, S, S, ri[:8], ci[:8], exact_bits=False)}
            extra[f"cfg3_{kind}"] = entry
            del A, B, C
    except Exception as e:
        extra["cfg3_error"] = repr(e)
    finally:
        qb.set_tensor_path(qb.TENSOR_AUTO)


def _parity_grid(torch, orc, A, lda, B, ldb, C, ldc, k, ri, ci, exact_bits, b_cols=None):
    """C[ri x ci] (device, row-major; ri relative to the local A / C block) against exact long-accumulator arithmetic (oracle/qoracle.c):
    only the sampled rows of A and columns of B leave the device.  alpha = 1, beta = 0 results.  b_cols(ci) -> (k, len(ci), 2) device
    tensor replaces the indexing of B (ranks that hold B as packed panels)."""
    from gpu_util import to_host
    from qblas_b200 import quad
    dev = C.device
    r_t, c_t = torch.as_tensor(ri, device=dev), torch.as_tensor(ci, device=dev)
    Ah = to_host(A.reshape(-1, lda, 2)[r_t][:, :k].contiguous().reshape(len(ri) * k, 2))
    Bsel = b_cols(ci) if b_cols is not None else B.reshape(-1, ldb, 2)[:k][:, c_t]
    Bh = to_host(Bsel.contiguous().reshape(k * len(ci), 2))
    got = to_host(C.reshape(-1, ldc, 2)[r_t][:, c_t].contiguous().reshape(len(ri) * len(ci), 2))
    idx = np.stack(np.meshgrid(np.arange(len(ri)), np.arange(len(ci)), indexing="ij"), axis=-1).reshape(-1, 2)
    exact, ratio, klass = orc.exact_dot_check("R", k, Ah, k, Bh, len(ci), idx, got)
    bad = int((ratio > 1.0).sum())
    out = {"checked_entries": int(idx.shape[0]), "contract_violations": bad, "worst_err_over_bound": float(ratio.max()),
           "against": "exact inner products (long accumulator, oracle/qoracle.c): |c^ - c| <= k u (|A||B|)_ij"}
    if exact_bits:
        out["bit_mismatches_vs_exact_rounded_once"] = int((~quad.same_bits(got, exact)).sum())
        out["against"] += " AND bit equality with the exact sum rounded once"
    return out


class ShardedGemm:
    """One row-sharded qgemm workload (BASELINE config 4 sharding): rank r owns rows [r m_loc, (r+1) m_loc) of A and C, B lives on rank 0.
    step() = qblas_b200.dist.qgemm_row_sharded (world > 1) or one device-resident qgemm (world = 1)."""

    def __init__(self, qb, torch, dist, rank, world, dev, M, n, k, kind, mode, args):
        from gpu_util import dev_random
        self.qb, self.torch, self.dist, self.rank, self.world, self.dev = qb, torch, dist, rank, world, dev
        self.M, self.n, self.k, self.m_loc, self.args, self.mode = M, n, k, M // world, args, mode
        m_loc = self.m_loc
        self.A = dev_random((m_loc * k,), kind, 100 + rank, dev)
        self.B = dev_random((k * n,), kind, 7, dev) if rank == 0 or world == 1 else torch.empty((k * n, 2), dtype=torch.int64, device=dev)
        Cinit = dev_random((m_loc * n,), kind, 9 + rank, dev)          # C_in of the local block (beta = 0 still reads it, level3.hpp:107)
        self.buf, self.gather, self.packed, self.panels, self.shared = None, "none", None, 0, None
        if world == 1:
            self.Cfull = Cinit
        else:
            fast = mode == qb.MODE_FAST
            from qblas_b200 import dist as qd
            self.qd = qd
            if args.gather == "fused" and fast:
                for cls, name in ((qd.SymmetricBuffer, "fused-multicast"), (qd.PeerBuffer, "fused-peer")):
                    if args.no_multicast and cls is qd.SymmetricBuffer:
                        continue
                    try:
                        self.buf = cls(M * n * 16)
                        self.gather = name if (cls is qd.PeerBuffer or self.buf.mc_ptr) else "fused-peer(symmetric memory, no multicast)"
                        break
                    except Exception as e:
                        print(f"[bench] {cls.__name__} unavailable ({e!r})", file=sys.stderr, flush=True)
                        self.buf = None
            if self.buf is not None:
                self.Cfull = self.buf.tensor.view(torch.int64).reshape(M * n, 2)
            else:
                self.Cfull = torch.empty((M * n, 2), dtype=torch.int64, device=dev)
                self.gather = "nccl"
            self.Cfull[rank * m_loc * n:(rank + 1) * m_loc * n].copy_(Cinit)
            del Cinit
            self.panels = args.panel_cols if (args.bcast == "panels" and fast and n >= 2 * args.panel_cols) else 0
            self.shared = None
            if self.panels:
                self.packed = torch.empty((k * n, 2), dtype=torch.int64, device=dev)
                if args.share_planes:
                    self.shared = qd.SharedPlanes()
        torch.cuda.empty_cache()     # hand freed blocks back to the driver: the library sizes its workspace from cudaMemGetInfo
        self.Cblk = self.Cfull[rank * m_loc * n:(rank + 1) * m_loc * n]

    def step(self):
        qb = self.qb
        if self.world == 1:
            qb.gemm("R", self.m_loc, self.n, self.k, 1.0, self.A, self.k, self.B, self.n, 0.0, self.Cblk, self.n)
            return
        self.qd.qgemm_row_sharded(self.M, self.n, self.k, 1.0, self.A, self.B, 0.0, self.Cfull, peers=self.buf, b_panels=self.panels, b_packed=self.packed,
                                  overlap_passes=(self.args.overlap if self.buf is None else 1), share_planes=self.shared)

    def describe(self):
        if self.world == 1:
            return "1 GPU, device resident"
        b = (f"B broadcast from rank 0 in packed column panels of {self.panels} columns DURING the product (qb_set_gemm_b_panels), column statistics first"
             if self.panels else "one NCCL broadcast of B before the product")
        if self.panels and self.shared is not None and self.shared.buf is not None:
            b += "; the residue planes of B are computed cooperatively (one column slice per rank) and exchanged through the NVSwitch multicast address of a symmetric buffer (dist.SharedPlanes)"
        g = {"nccl": "NCCL all_gather of the C blocks", "fused-peer": "fused gather: the reconstruction kernel stores each finished element into every peer's copy over NVLink (one store per peer)",
             "fused-multicast": "fused gather: the reconstruction kernel stores each finished element ONCE to the NVSwitch multicast address of torch symmetric memory"}.get(self.gather, self.gather)
        return f"C row-blocks of {self.m_loc} rows per GPU; {b}; {g}; completion barrier; all inside the timed region"

    def gather_check(self):
        """every rank must hold every rank's block: per-block checksums of the local C_full against the owners' own"""
        if self.world == 1:
            return 0
        torch, dist = self.torch, self.dist
        sums = self.Cfull.view(torch.int64).reshape(self.world, -1).sum(dim=1)            # wrapping int64 sums, one per block
        owners = [torch.empty(1, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
        dist.all_gather(owners, sums[self.rank:self.rank + 1].clone())
        return int((sums != torch.cat(owners)).sum().item())

    def close(self):
        self.Cblk = self.Cfull = None
        if self.buf is not None:
            self.buf.close()
        if getattr(self, "shared", None) is not None:
            self.shared.close()
        self.buf = self.packed = self.A = self.B = self.shared = None
        self.torch.cuda.empty_cache()


def _run_workload(w, steps, warmup, clocks=None):
    """W untimed steps, then exactly K timed steps bracketed by barrier + synchronize; CUDA events; max over ranks."""
    torch, dist, qb = w.torch, w.dist, w.qb

    def sync():
        torch.cuda.synchronize()
        if w.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        w.step()
    sync()
    if clocks is not None:
        clocks.start()
    l0 = qb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        w.step()
    ev1.record()
    sync()
    launches = qb.launch_count() - l0
    clk = clocks.stop() if clocks is not None else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=w.dev)
    if w.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    plan = qb.oz_last_stats() if w.mode == qb.MODE_FAST else None
    mma_ms, mma_launches = (qb.oz_last_mma_ms() if w.mode == qb.MODE_FAST else (0.0, 0))
    return {"ms_step": ms_step, "launches": launches, "clk": clk, "plan": plan, "mma_ms": mma_ms, "mma_launches": mma_launches,
            "gflops": 2.0 * w.M * w.n * w.k / (ms_step * 1e-3) / 1e9}


def _workload_parity(w, plan, nside):
    """nside x nside sampled entries of this rank's block (>= 4096 at nside = 64) against the oracle; summed over ranks by the caller"""
    import oracle_lib
    from gpu_util import to_host
    from qblas_b200 import quad
    qb, torch = w.qb, w.torch
    orc = oracle_lib.load_oracle()
    rng = np.random.default_rng(5 + w.rank)
    ri = np.sort(rng.choice(w.m_loc, min(nside, w.m_loc), replace=False)); ci = np.sort(rng.choice(w.n, min(nside, w.n), replace=False))
    if w.mode == qb.MODE_FAST:
        b_cols = None
        if w.world > 1 and w.panels:     # B itself is only valid on its owner; the ranks hold it as packed panels (panel j: k x cols, contiguous)
            pw = w.panels

            def b_cols(cs):
                out = []
                for c in cs:
                    c0 = (int(c) // pw) * pw
                    cols = min(pw, w.n - c0)
                    out.append(w.packed[c0 * w.k:(c0 + cols) * w.k].reshape(w.k, cols, 2)[:, int(c) - c0])
                return torch.stack(out, dim=1)
        par = _parity_grid(torch, orc, w.A, w.k, w.B, w.n, w.Cblk, w.n, w.k, ri, ci, exact_bits=bool(plan and plan["exact"]), b_cols=b_cols)
        mism = par["contract_violations"] + par.get("bit_mismatches_vs_exact_rounded_once", 0)
        return par["checked_entries"], mism, par["against"]
    # reference order: bit exact against the reference-order oracle (kc = 126), fewer entries (each is a chain of k rounded FMAs on the CPU)
    ri, ci = ri[:16], ci[:16]
    r_t, c_t = torch.as_tensor(ri, device=w.dev), torch.as_tensor(ci, device=w.dev)
    Ah = to_host(w.A.reshape(w.m_loc, w.k, 2)[r_t].contiguous().reshape(len(ri) * w.k, 2))
    Bh = to_host(w.B.reshape(w.k, w.n, 2)[:, c_t].contiguous().reshape(w.k * len(ci), 2))
    got = to_host(w.Cblk.reshape(w.m_loc, w.n, 2)[r_t][:, c_t].contiguous().reshape(-1, 2))
    idx = np.stack(np.meshgrid(np.arange(len(ri)), np.arange(len(ci)), indexing="ij"), axis=-1).reshape(-1, 2)
    exp = orc.gemm_sample("R", len(ri), len(ci), w.k, 1.0, Ah, w.k, Bh, len(ci), 0.0, None, len(ci), idx)
    return int(idx.shape[0]), int((~quad.same_bits(got, exp)).sum()), "oracle/qoracle.c, reference order (kc = 126), bit exact"


def _load_traffic(plan, m_unit, n_unit):
    """dram bytes per launch of k_oz_mma from the committed ncu --set full capture of the SAME launch shape (profiles/ncu_traffic.json), else None"""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        key = f"{m_unit}x{n_unit}x{plan['Kp']}x{plan['pairs']}"
        return tab.get(key)
    except Exception:
        return None


def own_arm(args, rank, world, local_rank):
    import torch
    import qblas_b200 as qb

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    qb.init()
    mode = qb.MODE_FAST if args.mode == "fast" else qb.MODE_REFERENCE
    qb.set_mode(mode)
    if args.host_slabs:
        qb.set_host_slabs(args.host_slabs)
    if args.unit:
        qb.set_tensor_unit(*[int(v) for v in args.unit.split(",")])
    if args.window:
        qb.set_tensor_window(args.window)
    S = args.size
    M, n, k = S * world, S, S
    strong = False
    if args.shape:   # BASELINE config 4 as the headline (tool runs): a FIXED global problem, C row-blocks of M / world rows per GPU
        M, n, k = (int(v) for v in args.shape.split(","))
        assert M % world == 0, "--shape M must be a multiple of the number of GPUs"
        strong = True
    dev = torch.device("cuda", local_rank)
    fast = mode == qb.MODE_FAST

    # ---- headline: inputs resident in HBM before the timed region (3 x 1 GiB at S=8192: far larger than the 126 MB L2)
    w = ShardedGemm(qb, torch, dist, rank, world, dev, M, n, k, args.dist, mode, args)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    res = _run_workload(w, args.steps, args.warmup, clocks)
    gather_bad = w.gather_check()
    plan = res["plan"]
    checked, mism, against = _workload_parity(w, plan, 64)
    describe = w.describe()
    m_loc = w.m_loc

    extra = {}
    roof = None
    if rank == 0:
        peaks, src = measured_peaks()
        if fast and plan and plan["pairs"] > 0:
            int8_ops = plan["pairs"] * 2.0 * m_loc * n * plan["Kp"]
            mma_ms = res["mma_ms"]
            tops = int8_ops / (mma_ms * 1e-3) / 1e12
            bf16_sus = float(peaks.get("bf16_tflops_sustained", 1400.0)); bf16_burst = float(peaks.get("bf16_tflops", 1590.0))
            peak = 2.0 * bf16_burst     # ONE rule: the burst figure (the higher one), whatever the launch length
            ur, uc = qb.get_tensor_unit()
            traffic = _load_traffic(plan, min(ur, m_loc), min(uc, n))
            roof = {"bound": "tensor", "kernel": "k_oz_mma<1> (tcgen05.mma kind::i8, TMA-fed, TMEM accumulators)", "achieved": tops, "peak": peak,
                    "unit": "TFLOP/s", "frac": tops / peak, "traffic": traffic,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this launch shape (profiles/ncu_traffic.json); null when no capture of this shape exists",
                    "peak_source": (f"2 x MEASURED_PEAKS.json bf16_tflops (burst {bf16_burst}; sustained {bf16_sus}) [{src}]: int8 dense issues at twice the bf16 rate on sm_100a. "
                                    "The kernel's own int8 rate on L2-resident operands is measured live in extra.int8_peak_microbench"),
                    "frac_of_2x_sustained": tops / (2.0 * bf16_sus),
                    "algorithmic": (f"one binary128 flop = {plan['pairs']} int8 ops: row/column block fixed point (W_A = {plan['WA']}, W_B = {plan['WB']} bits), one int8 GEMM per "
                                    f"modulus, {plan['pairs']} pairwise coprime moduli <= 256 (product > 2 k 2^(W_A+W_B)), exact CRT reconstruction: {plan['pairs']} x 2*m*n*Kp = "
                                    f"{int8_ops:.4g} int8 ops per qgemm in {res['mma_launches']} launches of k_oz_mma ({plan['row_passes']} row passes x {plan['panels']} column panels), "
                                    f"{mma_ms:.2f} ms summed (CUDA events on the launching stream, last timed step); whole step {res['ms_step']:.2f} ms"),
                    "kernel_ms": mma_ms, "kernel_share_of_step": mma_ms / res["ms_step"], "whole_step_frac": int8_ops / (res["ms_step"] * 1e-3) / 1e12 / peak,
                    "plan": plan, "binary128_gflops_in_kernel": 2.0 * m_loc * n * k / (mma_ms * 1e-3) / 1e9}
        else:
            pk = _int_issue_peak(qb, torch, dev)
            kern_gflops = 2.0 * m_loc * n * k / (res["ms_step"] * 1e-3) / 1e9
            floor_gf = 2.0 * 32.0 * 148 * 4 * 1.965 / 69.0   # the step as nothing but the 16 IMAD.WIDE of its exact product: 69 dispatch clocks per warp-qFMA
            roof = {"bound": "integer dispatch (sum of the dispatch clocks of the qFMA's instruction stream; not hbm, not tensor)", "kernel": "k_gemm_nb",
                    "achieved": kern_gflops, "peak": floor_gf,
                    "unit": "GFLOP/s (binary128)", "frac": kern_gflops / floor_gf, "traffic": None,
                    "peak_source": "16 IMAD.WIDE x 4.3 dispatch clocks per warp-qFMA and SM sub-partition at 1.965 GHz (tools/exp/mb_pipes.cu, tools/sass_cost.py); the register-resident microbenchmark of the first-version step is in register_resident_microbench_gflops",
                    "register_resident_microbench_gflops": pk, "frac_of_that_microbench": kern_gflops / pk,
                    "algorithmic": f"2*m*n*k = {2.0 * m_loc * n * k:.4g} binary128 flops per launch; avg step {res['ms_step']:.2f} ms (CUDA events)"}

    # ---- e2e: reference-named C entry point with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not strong:
        from gpu_util import to_host  # noqa: F401
        hA = torch.empty((m_loc * k, 2), dtype=torch.int64).pin_memory(); hA.copy_(w.A)
        hB = torch.empty((k * n, 2), dtype=torch.int64).pin_memory()
        if world > 1:
            Bfull = w.B if rank == 0 else torch.empty((k * n, 2), dtype=torch.int64, device=dev)
            dist.broadcast(Bfull.view(torch.uint8), src=0)
            hB.copy_(Bfull); del Bfull
        else:
            hB.copy_(w.B)
        hC = torch.empty((m_loc * n, 2), dtype=torch.int64).pin_memory(); hC.copy_(w.Cblk)
    w.close()
    del w
    if not strong:
        nA, nB, nC = (x.numpy().view(np.uint64) for x in (hA, hB, hC))
        e2e_steps = max(1, min(args.steps, 3))
        qb.quadblas_qgemm("R", "N", "N", m_loc, n, k, 1.0, nA, k, nB, n, 0.0, nC, n)  # warm (allocates staging)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            qb.quadblas_qgemm("R", "N", "N", m_loc, n, k, 1.0, nA, k, nB, n, 0.0, nC, n)  # synchronous: returns with C on the host
        t1 = time.perf_counter()
        te = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # beta = 0 and contiguous rows of C: the library classifies C_in on the host and uploads one class byte per element instead of 16
        # (qb_set_beta0_classes; beta * C_in keeps the reference's bits), so that is what crosses PCIe
        cls = bool(qb.get_beta0_classes())
        e2e = {"value": 2.0 * M * n * k / float(te.item()) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 16 * (m_loc * k + k * n) + (1 if cls else 16) * m_loc * n,
               "c_in": "one class byte per element (beta = 0: finite sign / Inf-NaN), read in full by host threads" if cls else "16 bytes per element",
               "d2h_bytes_per_step": 16 * m_loc * n, "ms_per_step": float(te.item()) * 1e3, "steps": e2e_steps,
               "api": "quadblas_qgemm (reference C ABI), pinned host buffers, synchronous" +
                      ("" if world == 1 else "; N > 1: every rank multiplies ITS row block from its own host buffers at the same time (no broadcast / gather: "
                                             "NOT the same job as `value`; the ranks share the host's memory and PCIe root)"),
               "plan": qb.oz_last_stats() if fast else None}
        del hA, hB, hC, nA, nB, nC

    # ---- secondary figures (rank 0 computes the single-GPU ones; the multi-GPU ones are collective)
    if not args.no_extra:
        if rank == 0 and not strong:
            if fast:
                try:
                    tops8, ms8 = _int8_peak(qb, torch, dev)
                    extra["int8_peak_microbench"] = {"tops": tops8, "ms_per_launch": ms8,
                                                     "what": "k_oz_mma<0> (same TMA + tcgen05 kind::i8 pipeline, int32 output), 64 plane pairs of 2048 x 2048 x 2048 per launch, operands L2 resident, 20 launches back to back"}
                    if roof is not None and roof.get("bound") == "tensor":
                        roof["frac_of_measured_int8_peak"] = roof["achieved"] / tops8
                except Exception as e:
                    extra["int8_peak_microbench"] = {"error": repr(e)}
        if world == 1 and not strong:
            _secondary(qb, torch, dev, args, S, mode, extra)
            if fast and S >= 4096:
                _cfg3(qb, torch, dev, S, extra)
        if fast and not strong and args.cfg4:
            try:
                extra["cfg4_strong"] = _cfg4(qb, torch, dist, rank, world, dev, args)
            except Exception as e:
                extra["cfg4_strong"] = {"error": repr(e)}
                torch.cuda.empty_cache()
        if world > 1 and not strong:
            try:
                extra.update(_mgpu_level12(qb, torch, dist, rank, world, dev))
            except Exception as e:
                extra["mgpu_error"] = repr(e)

    _print_line(args, rank, world, dist, torch, dev, qb, mode, plan, res, M, n, k, m_loc, e2e, roof, checked, mism, against, gather_bad, describe, extra, strong)


def _cfg4(qb, torch, dist, rank, world, dev, args):
    """BASELINE config 4: quadblas_qgemm 32768^3 as a fixed global problem, C row-blocks of 32768 / N rows (strong scaling)."""
    M = n = k = args.cfg4
    w = ShardedGemm(qb, torch, dist, rank, world, dev, M, n, k, "D113", qb.MODE_FAST, args)
    res = _run_workload(w, 3, 1)
    bad = w.gather_check()
    checked, mism, against = _workload_parity(w, res["plan"], 32)
    t = torch.tensor([checked, mism, bad], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    desc = w.describe()
    m_loc = w.m_loc
    w.close()
    plan = res["plan"]
    peaks, _ = measured_peaks()
    peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
    int8_ops = plan["pairs"] * 2.0 * m_loc * n * plan["Kp"]
    return {"workload": f"quadblas_qgemm row-major {M}x{n}x{k} alpha=1 beta=0, D113, fixed global problem ({desc})", "n_gpus": world, "steps": 3, "warmup": 1,
            "ms_per_step": res["ms_step"], "gflops": res["gflops"], "scaling": "strong",
            "tensor_kernel_ms": res["mma_ms"], "tensor_kernel_tops_per_gpu": int8_ops / (res["mma_ms"] * 1e-3) / 1e12,
            "whole_step_tensor_frac_per_gpu": int8_ops / (res["ms_step"] * 1e-3) / 1e12 / peak, "tensor_peak_tops": peak, "plan": plan,
            "parity": {"checked_entries": int(t[0].item()), "mismatches": int(t[1].item()), "against": against,
                       "gathered_blocks_checked": world * world if world > 1 else 0, "gathered_blocks_wrong": int(t[2].item())}}


def _mgpu_level12(qb, torch, dist, rank, world, dev):
    """Row-sharded qgemv (32768^2) and range-sharded qdot (n = 10^8) on N GPUs through qblas_b200.dist, fast mode: each rank streams
    its shard from HBM; x is broadcast and y gathered / the 16-byte partials are all-gathered and folded in rank order on every rank.
    Bitwise against the single-GPU call on rank 0 for qgemv (row sharding changes no element's arithmetic); qdot against exact arithmetic."""
    from gpu_util import dev_random, to_host
    from qblas_b200 import dist as qd
    import oracle_lib
    peaks, src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {}
    qb.set_mode(qb.MODE_FAST)
    # ---- qgemv
    m = n = 32768
    lo, hi = qd.row_block(m, world, rank)
    gen = torch.Generator(device=dev); gen.manual_seed(77)
    Ablk = dev_random(((hi - lo) * n,), "D113", 500 + rank, dev)
    x = dev_random((n,), "D113", 12, dev) if rank == 0 else torch.zeros((n, 2), dtype=torch.int64, device=dev)
    y = torch.zeros((m, 2), dtype=torch.int64, device=dev)

    def gv():
        qd.qgemv_row_sharded(m, n, 1.0, Ablk, x, 0.0, y)
    gv(); gv()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = _time_events(gv, 5)
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    # parity: the owner's rows recomputed by ONE GPU call on the same rows must give the same bits (and every rank holds them)
    y1 = torch.zeros((hi - lo, 2), dtype=torch.int64, device=dev)
    qb.gemv("R", hi - lo, n, 1.0, Ablk, n, x, 1, 0.0, y1, 1, m_total=m)    # these rows as ONE GPU computing the whole m x n qgemv produces them
    same = int((y[lo:hi] == y1).all().item())
    chk = y.view(torch.int64).sum().reshape(1); allc = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(allc, chk)
    same_all = int(all(int(c.item()) == int(chk.item()) for c in allc))
    byt = 16.0 * (m * n + n + 2 * m)
    out["mgpu_qgemv"] = {"workload": f"quadblas_qgemv R/N {m}x{n} fast mode, row blocks of {hi - lo} rows per GPU, broadcast(x) + gather(y) (qblas_b200.dist.qgemv_row_sharded)",
                         "n_gpus": world, "ms": ms, "gbs_total": byt / ms / 1e6, "gbs_per_gpu": byt / world / ms / 1e6, "frac_of_hbm_per_gpu": byt / world / ms / 1e6 / hbm,
                         "bitwise_equal_to_single_gpu_rows": bool(same), "all_ranks_hold_same_y": bool(same_all), "peak_source": src}
    del Ablk, y, y1
    # ---- qdot
    nd = 100_000_000
    lo, hi = qd.dot_shard_range(nd, 1, world, rank, False)
    xs = dev_random((hi - lo,), "D113", 600 + rank, dev); ys = dev_random((hi - lo,), "D113", 700 + rank, dev)
    res = torch.zeros((1, 2), dtype=torch.int64, device=dev)

    def dt():
        qd.qdot_sharded(nd, xs, ys, 1, res, reference_order=False)
    dt(); dt()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = _time_events(dt, 5)
    t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    allr = [torch.empty_like(res) for _ in range(world)]; dist.all_gather(allr, res)
    same_all = int(all(bool((r == res).all().item()) for r in allr))
    # contract of the local partial against exact arithmetic on a bounded prefix of the shard (the CPU check is O(n))
    npre = min(hi - lo, 2_000_000)
    part = torch.zeros((1, 2), dtype=torch.int64, device=dev)
    qb.dot(npre, xs, 1, ys, 1, part)
    orc = oracle_lib.load_oracle()
    _, ratio, klass = orc.exact_dot_check("R", npre, to_host(xs[:npre]), npre, to_host(ys[:npre]), 1, np.array([[0, 0]]), to_host(part).reshape(1, 2))
    out["mgpu_qdot"] = {"workload": f"qdot n={nd} fast mode, contiguous ranges of {hi - lo} per GPU, NCCL all_gather of the 16-byte partials, fold in rank order on every rank (qblas_b200.dist.qdot_sharded)",
                        "n_gpus": world, "ms": ms, "gbs_total": 32.0 * nd / ms / 1e6, "gbs_per_gpu": 32.0 * nd / world / ms / 1e6, "frac_of_hbm_per_gpu": 32.0 * nd / world / ms / 1e6 / hbm,
                        "all_ranks_hold_same_result": bool(same_all), "local_partial_err_over_bound": float(ratio[0]), "peak_source": src}
    qb.set_mode(qb.MODE_REFERENCE)
    return out


def _mode_text(plan):
    if plan and plan["pairs"] > 0:
        t = (f"fast: exact int8 residue planes on tcgen05 ({plan['pairs']} moduli, one GEMM each) + Chinese-remainder reconstruction: ")
        return t + ("inner products exact, rounded once" if plan["exact"] else
                    f"windows capped at {plan['WA']} / {plan['WB']} bits (spans {plan['WA_span']} / {plan['WB_span']}), every element tested, {plan['flagged']} recomputed by the fix-up: inside gamma_k (|A||B|)")
    return "fast: integer-limb kernel (the tensor path declined)"


def _print_line(args, rank, world, dist, torch, dev, qb, mode, plan, res, M, n, k, m_loc, e2e, roof, checked, mism, against, gather_bad, describe, extra, strong):
    mism_t = torch.tensor([checked, mism, gather_bad], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(mism_t)
    if rank == 0:
        cpu = None
        if not strong:
            try:
                cpu = time_reference_gemm(12.0)
                cpu = {k2: cpu[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        fast = mode == qb.MODE_FAST
        line = {
            "metric": "binary128 qgemm GFLOPS", "value": res["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "binary128 (exact int8 residues on the tensor cores, Chinese-remainder recombination, one rounding)" if fast and plan and plan["pairs"] > 0
                     else "binary128 (software, u32 integer limbs)",
            "data": "synthetic",
            "config": {"workload": f"quadblas_qgemm row-major {M}x{n}x{k} alpha=1 beta=0 ({describe if world > 1 else ('BASELINE config 3, 1xB200' if not strong else 'BASELINE config 4 shape on 1 GPU')})",
                       "mode": _mode_text(plan) if fast else "reference-order (bit exact, kc=126), integer-limb kernel",
                       "inputs": {"D113": "D113: full 113-bit random mantissas, device resident", "D53": "D53: doubles U(-1,1) cast to quad (the reference's own benchmark distribution)",
                                  "Dexp": "Dexp: D113 x 2^U{-40..40}"}[args.dist],
                       "l2": "inputs (3 x 1 GiB at 8192^3) and the int8 planes (GBs) exceed the 126 MB L2; no flush needed", "parallelism": f"row-block x{world}"},
            "e2e": e2e, "gpu_launches": int(res["launches"]), "clocks": res["clk"], "roofline": roof, "cpu_baseline": cpu,
            "parity": {"checked_entries": int(mism_t[0].item()), "mismatches": int(mism_t[1].item()), "against": against,
                       "gathered_blocks_checked": world * world if world > 1 else 0, "gathered_blocks_wrong": int(mism_t[2].item())},
            "extra": extra,
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
    ap.add_argument("--overlap", type=int, default=4, help="N > 1, --gather nccl: row passes whose all-gathers overlap the next pass (1 = one all-gather after the qgemm)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N > 1: C blocks reach the other ranks by stores from the kernel that finishes them (fused) or by NCCL all-gather")
    ap.add_argument("--no-multicast", action="store_true", help="N > 1, --gather fused: CUDA IPC peer buffers (one store per peer) instead of torch symmetric memory + NVSwitch multicast")
    ap.add_argument("--bcast", default="panels", choices=["panels", "whole"], help="N > 1: B is broadcast in column panels during the product, or as a whole before it")
    ap.add_argument("--share-planes", action="store_true", help="N > 1, --bcast panels: the ranks reduce one column slice of B each and exchange the residue planes over the multicast fabric (dist.SharedPlanes; measured slower than every rank reducing all of B: profiles/r2_shared_planes_probe_n8.log)")
    ap.add_argument("--panel-cols", type=int, default=2048, help="N > 1, --bcast panels: columns per panel (multiple of 256)")
    ap.add_argument("--host-slabs", type=int, default=0, help="e2e: slabs of the pipelined all-host qgemm (library default 8)")
    ap.add_argument("--unit", default=None, help="tensor path pipeline unit rows,cols (library default 2048,4096)")
    ap.add_argument("--window", type=int, default=0, help="tensor path window budget in bits (library default 144)")
    ap.add_argument("--cfg4", type=int, default=32768, help="side of the fixed global problem of extra.cfg4_strong (0 = skip)")
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
