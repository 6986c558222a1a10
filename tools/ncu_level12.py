"""Tiny driver for `ncu --set full` captures of the fast-mode level-1/2 kernels (one launch each after a warm-up):
qgemv R/N and C/N 16384^2, qdot / qnrm2 n = 2e7.  Run under ncu with -k regex:'k_gemv_.*wide|k_dot_wide_l1'."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_FAST)
m = 16384; n = 20_000_000
A = dev_random((m * m,), "D113", 11, dev); x = dev_random((m,), "D113", 12, dev); y = dev_random((m,), "D113", 13, dev)
xd = dev_random((n,), "D113", 14, dev); yd = dev_random((n,), "D113", 15, dev)
res = torch.zeros((1, 2), dtype=torch.int64, device=dev)
for _ in range(2):
    qb.gemv("R", m, m, 1.0, A, m, x, 1, 0.0, y, 1)
    qb.gemv("C", m, m, 1.0, A, m, x, 1, 0.0, y, 1)
    qb.dot(n, xd, 1, yd, 1, res)
    qb.nrm2(n, xd, 1, res)
torch.cuda.synchronize()
qb.set_mode(qb.MODE_REFERENCE)
