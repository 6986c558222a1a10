"""Tiny driver for ncu captures of the fast-mode qgemm (one warm-up call + one profiled-size call, device resident):
usage: ncu ... python tools/ncu_qgemm.py [size=8192] [dist=D113|D53|Dexp] [calls=2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kind = sys.argv[2] if len(sys.argv) > 2 else "D113"
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_FAST)
A = dev_random((S * S,), kind, 100, dev); B = dev_random((S * S,), kind, 7, dev); C = dev_random((S * S,), kind, 9, dev)   # bench.py's rank-0 inputs: the same plan
for _ in range(calls):
    qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
torch.cuda.synchronize()
print(qb.oz_last_stats())
qb.set_mode(qb.MODE_REFERENCE)
