"""Tiny driver for an ncu capture of the reference-order integer-limb qgemm (k_gemm): usage: ncu ... python tools/ncu_kgemm.py [size=2048]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_REFERENCE)
A = dev_random((S * S,), "D113", 1, dev); B = dev_random((S * S,), "D113", 2, dev); C = dev_random((S * S,), "D113", 3, dev)
for _ in range(2):
    qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
torch.cuda.synchronize()
