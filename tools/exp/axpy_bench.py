"""qaxpy n = 10^8 against HBM (48 B per element).  Development tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
qb.init()
for n in (10 ** 7, 10 ** 8):
    x = dev_random((n,), "D113", 5); y = dev_random((n,), "D113", 6)
    for _ in range(2):
        qb.axpy(n, 1.5, x, 1, y, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); qb.axpy(n, 1.5, x, 1, y, 1); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(f"axpy n={n}: {ms:.4f} ms  {48.0 * n / ms * 1e-9:.3f} TB/s  {n / ms * 1e-6:.1f} G qFMA/s", flush=True)
