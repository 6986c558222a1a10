"""What one rank of BASELINE config 4 computes at 8 GPUs: a 4096 x 32768 x 32768 qgemm (device resident), for several unit shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
m, n, k = 4096, 32768, 32768
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_FAST)
A = dev_random((m * k,), "D113", 1, dev); B = dev_random((k * n,), "D113", 2, dev); C = dev_random((m * n,), "D113", 3, dev)
torch.cuda.empty_cache()
for sh in (sys.argv[1] if len(sys.argv) > 1 else "2048x2048,4096x2048,2048x4096,4096x4096,4096x1024").split(","):
    ur, uc = (int(v) for v in sh.split("x"))
    qb.set_tensor_unit(ur, uc)
    qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        qb.gemm("R", m, n, k, 1.0, A, k, B, n, 0.0, C, n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    mma, nl = qb.oz_last_mma_ms()
    tl = qb.oz_last_mma_timeline()
    gaps = [tl[i + 1][0] - (tl[i][0] + tl[i][1]) for i in range(len(tl) - 1)]
    st = qb.oz_last_stats()
    print(f"unit {ur}x{uc}: {ms:.1f} ms/call ({2 * m * n * k / ms / 1e9:.1f} TFLOP/s), mma sum {mma:.1f} ms in {nl} launches, first launch at +0, span {tl[-1][0] + tl[-1][1]:.1f} ms, "
          f"gaps total {sum(gaps):.2f} ms (max {max(gaps):.2f}), moduli {st['pairs']}, ws {st['ws_bytes'] / 2**30:.1f} GiB", flush=True)
qb.set_mode(qb.MODE_REFERENCE)
