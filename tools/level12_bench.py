"""BASELINE config 2 on one B200: quadblas_qdot n = 1e5..1e8, qnrm2, quadblas_qgemv 1500^2..32768^2 (row- and
col-major), device-resident, CUDA events, >= 3 warm-ups, median of `--reps`; fast mode (window accumulator
and the older rounded-FMA-chain variant) and reference-order mode, each as GB/s of ALGORITHMIC bytes
(dot 32n, nrm2 16n, gemv 16(mn + n + 2m)) against MEASURED_PEAKS.json's HBM figure, plus the
register-resident microbenchmarks that give each accumulate's integer-issue ceiling.
Prints one JSON object; `--out` also writes it to a file."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--max-gemv", type=int, default=32768)
    ap.add_argument("--max-dot", type=int, default=100_000_000)
    ap.add_argument("--dist", default="D113")
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-reference", action="store_true")
    args = ap.parse_args()
    import torch
    import qblas_b200 as qb
    from gpu_util import dev_random
    dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        hbm = 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timeit(fn, flush_l2):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.reps):
            if flush_l2:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    out = {"hbm_peak_gbs": hbm, "dist": args.dist, "reps": args.reps, "dot": [], "nrm2": [], "gemv": [], "microbench": {}}
    sink = torch.zeros((4, 2), dtype=torch.int64, device=dev)
    for name, variants in (("rounded_fma", ((104, 256), (4, 256), (2, 256))), ("window", ((1004, 256), (1002, 256), (1001, 256), (1004, 128)))):
        best = 0.0
        for v, th in variants:
            qb.fma_microbench(v, 148 * 4, th, 64, sink)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); nf = qb.fma_microbench(v, 148 * 4, th, 2000, sink); b.record(); torch.cuda.synchronize()
            r = nf / (a.elapsed_time(b) * 1e-3)
            out["microbench"][f"{name}_{v}_{th}"] = r / 1e9
            best = max(best, r)
        out["microbench"][name + "_best_Gacc_per_s"] = best / 1e9
    modes = [("fast_window", qb.MODE_FAST, 1), ("fast_rounded_chain", qb.MODE_FAST, 0)]
    if not args.skip_reference:
        modes.append(("reference_order", qb.MODE_REFERENCE, 1))
    res = torch.zeros((1, 2), dtype=torch.int64, device=dev)
    nmax = args.max_dot
    xd = dev_random((nmax,), args.dist, 14, dev); yd = dev_random((nmax,), args.dist, 15, dev)
    for n in (100_000, 1_000_000, 10_000_000, 100_000_000):
        if n > nmax:
            continue
        for name, md, fv in modes:
            qb.set_mode(md); qb.set_fast_variant(fv)
            if md == qb.MODE_REFERENCE:
                qb.quadblas_set_num_threads(4096)
            small = 32 * n < (100 << 20)
            ms = timeit(lambda: qb.dot(n, xd, 1, yd, 1, res), small)
            out["dot"].append({"n": n, "mode": name, "ms": ms, "gbs": 32.0 * n / ms / 1e6, "frac_hbm": 32.0 * n / ms / 1e6 / hbm, "l2_flushed": small})
            ms = timeit(lambda: qb.nrm2(n, xd, 1, res), small)
            out["nrm2"].append({"n": n, "mode": name, "ms": ms, "gbs": 16.0 * n / ms / 1e6, "frac_hbm": 16.0 * n / ms / 1e6 / hbm, "l2_flushed": small})
            qb.quadblas_set_num_threads(0)
    del xd, yd
    mmax = args.max_gemv
    Av = dev_random((mmax * mmax,), args.dist, 11, dev); xv = dev_random((mmax,), args.dist, 12, dev); yv = dev_random((mmax,), args.dist, 13, dev)
    for m in (1500, 4096, 8192, 16384, 32768):
        if m > mmax:
            continue
        for layout in "RC":
            for name, md, fv in modes:
                qb.set_mode(md); qb.set_fast_variant(fv)
                byt = 16.0 * (m * m + m + 2 * m)
                small = byt < (100 << 20)
                ms = timeit(lambda: qb.gemv(layout, m, m, 1.0, Av, m, xv, 1, 0.0, yv, 1), small)
                out["gemv"].append({"m": m, "layout": layout, "mode": name, "ms": ms, "gbs": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / hbm,
                                    "gflops": 2.0 * m * m / ms / 1e6, "l2_flushed": small})
    qb.set_mode(qb.MODE_REFERENCE); qb.set_fast_variant(1)
    s = json.dumps(out)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
