"""qgemv fast mode, row-major: sliced FP64 accumulate (variant 2) against the window accumulator (variant 1).  Development tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random

qb.init(); qb.set_mode(qb.MODE_FAST)
for n in (8192, 32768):
    A = dev_random((n * n,), "D113", 5); x = dev_random((n,), "D113", 6); y = dev_random((n,), "D113", 7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for var, lay in ((1, "R"), (2, "R"), (1, "C"), (2, "C")):
        qb.set_fast_variant(var)
        for _ in range(3):
            qb.gemv(lay, n, n, 1.0, A, n, x, 1, 0.0, y, 1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); qb.gemv(lay, n, n, 1.0, A, n, x, 1, 0.0, y, 1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        by = 16.0 * (n * n + n + 2 * n)
        print(f"n={n} {lay} variant {var}: {ms:.3f} ms  {by / ms * 1e-9:.3f} TB/s  declined {qb.gemv_last_declined()}", flush=True)
qb.set_fast_variant(2); qb.set_mode(qb.MODE_REFERENCE)
