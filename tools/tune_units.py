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
ramps = sys.argv[4] if len(sys.argv) > 4 else "0x0"
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_FAST)
A = dev_random((S * S,), kind, 1, dev); B = dev_random((S * S,), kind, 2, dev); C = dev_random((S * S,), kind, 3, dev)
qb.set_tensor_unit(2048, 2048)
for _ in range(10):      # bring the clocks up before the first measured shape
    qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
for sh, rp in [(a, b) for a in shapes.split(",") for b in ramps.split(",")]:
    ur, uc = (int(v) for v in sh.split("x"))
    qb.set_tensor_unit(ur, uc)
    qb.set_tensor_ramp(*[int(v) for v in rp.split("x")])
    for _ in range(3):
        qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
    torch.cuda.synchronize()
    NIT = 20
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(NIT + 1)]
    evs[0].record()
    import time
    host = []
    for i in range(NIT):
        t0 = time.perf_counter()
        qb.gemm("R", S, S, S, 1.0, A, S, B, S, 0.0, C, S)
        host.append((time.perf_counter() - t0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(NIT)]
    ms = sum(per) / NIT
    print("   per-call ms (device):", [round(v, 1) for v in per], " host call ms:", [round(v, 1) for v in host])
    mma, nl = qb.oz_last_mma_ms()
    tl = qb.oz_last_mma_timeline()
    span = tl[-1][0] + tl[-1][1]
    gaps = [round(tl[i + 1][0] - (tl[i][0] + tl[i][1]), 3) for i in range(len(tl) - 1)]
    st = qb.oz_last_stats()
    print(f"unit {ur}x{uc} ramp {rp}: {ms:.2f} ms/call ({2 * S**3 / ms / 1e9:.1f} TFLOP/s), mma sum {mma:.2f} ms in {nl} launches over a span of {span:.2f} ms, moduli {st['pairs']}, flagged {st['flagged']}", flush=True)
    print("   launch ms:", [round(d, 2) for _, d in tl][:40])
    print("   gaps   ms:", gaps[:40], flush=True)
qb.set_mode(qb.MODE_REFERENCE)
