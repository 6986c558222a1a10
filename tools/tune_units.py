"""Tuning aid for the tensor-path pipeline: times the device-resident fast-mode qgemm for several unit shapes and prints the tensor
kernel's launch timeline (gaps = pipeline bubbles).  usage: python tools/tune_units.py [size=8192] [dist=D113] [shapes=2048x2048,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kind = sys.argv[2] if len(sys.argv) > 2 else "D113"
shapes = sys.argv[3] if len(sys.argv) > 3 else "2048x2048,2048x8192,2048x4096,4096x2048,1024x2048,4096x4096,8192x8192"
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_FAST)
A = dev_random((S * S,), kind, 1, dev); B = dev_random((S * S,), kind, 2, dev); C = dev_random((S * S,), kind, 3, dev)
for sh in shapes.split(","):
    ur, uc = (int(v) for v in sh.split("x"))
    qb.set_tensor_unit(ur, uc)
    for _ in range(3):
        qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    mma, nl = qb.oz_last_mma_ms()
    tl = qb.oz_last_mma_timeline()
    span = tl[-1][0] + tl[-1][1]
    gaps = [round(tl[i + 1][0] - (tl[i][0] + tl[i][1]), 3) for i in range(len(tl) - 1)]
    st = qb.oz_last_stats()
    print(f"unit {ur}x{uc}: {ms:.2f} ms/call ({2 * S**3 / ms / 1e9:.1f} TFLOP/s), mma sum {mma:.2f} ms in {nl} launches over a span of {span:.2f} ms, moduli {st['pairs']}, flagged {st['flagged']}", flush=True)
    print("   launch ms:", [round(d, 2) for _, d in tl][:40])
    print("   gaps   ms:", gaps[:40], flush=True)
qb.set_mode(qb.MODE_REFERENCE)
