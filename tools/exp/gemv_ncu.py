"""one fast-mode row-major qgemv (sliced FP64 kernel) for an ncu capture.  Development tool."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
qb.init(); qb.set_mode(qb.MODE_FAST); qb.set_fast_variant(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
lay = sys.argv[3] if len(sys.argv) > 3 else "R"
A = dev_random((n * n,), "D113", 5); x = dev_random((n,), "D113", 6); y = dev_random((n,), "D113", 7)
for _ in range(2):
    qb.gemv(lay, n, n, 1.0, A, n, x, 1, 0.0, y, 1)
torch.cuda.synchronize()
