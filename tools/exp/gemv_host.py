"""quadblas_qgemv 16384^2 (4 GiB) fast mode from pageable numpy arrays.  Development tool."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qblas_b200 as qb
from gpu_util import dev_random, to_host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
qb.init(); qb.set_mode(qb.MODE_FAST)
A = to_host(dev_random((S * S,), "D113", 1)); x = to_host(dev_random((S,), "D113", 2)); y = to_host(dev_random((S,), "D113", 3))
qb.quadblas_qgemv("R", "N", S, S, 1.0, A, S, x, 1, 0.0, y, 1)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); qb.quadblas_qgemv("R", "N", S, S, 1.0, A, S, x, 1, 0.0, y, 1); ts.append(time.perf_counter() - t0)
print(f"qgemv {S}^2 pageable: {min(ts) * 1e3:.1f} ms  {16.0 * S * S / min(ts) / 1e9:.1f} GB/s of A", flush=True)
