import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
n = 10 ** 8
qb.init(); qb.set_mode(qb.MODE_FAST)
out = torch.zeros(2, dtype=torch.int64, device="cuda")
x = dev_random((n,), "D113", 5); y = dev_random((n,), "D113", 6)
for _ in range(2):
    qb.dot(n, x, 1, y, 1, out)
torch.cuda.synchronize()
