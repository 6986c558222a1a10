"""qnrm2 / qdot fast mode: sliced FP64 sum of squares (variant 2) against the window accumulator (variant 1).  Development tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
qb.init(); qb.set_mode(qb.MODE_FAST)
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for n in (4 * 10 ** 6, 10 ** 7, 3 * 10 ** 7):
    x = dev_random((n,), "D113", 5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for var in (2, 3):
        qb.set_fast_variant(var)
        for _ in range(3):
            qb.nrm2(n, x, 1, out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); qb.nrm2(n, x, 1, out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        print(f"nrm2 n={n} variant {var}: {ms:.4f} ms  {16.0 * n / ms * 1e-9:.3f} TB/s", flush=True)
qb.set_fast_variant(2); qb.set_mode(qb.MODE_REFERENCE)
