"""qdot fast mode: sliced FP64 dot of two vectors (variant 2/3) against the window accumulator (variant 1).  Development tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
qb.init(); qb.set_mode(qb.MODE_FAST)
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for n in (10 ** 7, 10 ** 8):
    x = dev_random((n,), "D113", 5); y = dev_random((n,), "D113", 6)
    for var in (1, 3, 1, 3):
        qb.set_fast_variant(var)
        for _ in range(3):
            qb.dot(n, x, 1, y, 1, out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); qb.dot(n, x, 1, y, 1, out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        print(f"dot n={n} variant {var}: {ms:.4f} ms  {32.0 * n / ms * 1e-9:.3f} TB/s", flush=True)
qb.set_fast_variant(2); qb.set_mode(qb.MODE_REFERENCE)
