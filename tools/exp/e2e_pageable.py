"""quadblas_qgemm 8192^3 fast mode through the reference C ABI with PAGEABLE host buffers (what std::vector / numpy callers pass),
against page-locked ones.  Development tool."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qblas_b200 as qb
from gpu_util import dev_random, to_host
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
qb.init(); qb.set_mode(qb.MODE_FAST)
A = to_host(dev_random((S * S,), "D113", 1)); B = to_host(dev_random((S * S,), "D113", 2)); C = to_host(dev_random((S * S,), "D113", 3))
def run(a, b, c, tag):
    qb.quadblas_qgemm("R", "N", "N", S, S, S, 1.0, a, S, b, S, 0.0, c, S)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); qb.quadblas_qgemm("R", "N", "N", S, S, S, 1.0, a, S, b, S, 0.0, c, S); ts.append(time.perf_counter() - t0)
    print(f"{tag}: {min(ts) * 1e3:.1f} ms  {2.0 * S ** 3 / min(ts) / 1e12:.2f} TFLOP/s", flush=True)
run(A, B, C, "pageable numpy")
pA = torch.from_numpy(A.view(np.int64)).pin_memory(); pB = torch.from_numpy(B.view(np.int64)).pin_memory(); pC = torch.from_numpy(C.view(np.int64)).pin_memory()
run(pA.numpy().view(np.uint64), pB.numpy().view(np.uint64), pC.numpy().view(np.uint64), "pinned")
