"""fast qdot at several sizes (one launch each after a warm-up): for an ncu launch list of kernel durations against n, and the same
sizes timed with events over 20 back-to-back calls."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import qblas_b200 as qb
from gpu_util import dev_random
qb.init(); qb.set_mode(qb.MODE_FAST)
out = torch.zeros(2, dtype=torch.int64, device="cuda")
N = 10 ** 8
x = dev_random((N,), "D113", 5); y = dev_random((N,), "D113", 6)
sizes = (1 << 19, 10 ** 6, 3 * 10 ** 6, 10 ** 7, 3 * 10 ** 7, 10 ** 8)
for n in sizes:
    qb.dot(n, x, 1, y, 1, out)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    sys.exit(0)
for n in sizes:
    reps = 20
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        qb.dot(n, x, 1, y, 1, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"dot n={n}: {ms * 1e3:.1f} us per call over {reps} back-to-back calls  {32.0 * n / ms * 1e-9:.3f} TB/s", flush=True)
