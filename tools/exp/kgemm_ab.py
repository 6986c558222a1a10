"""Reference-order qgemm, the two kernels side by side (QBLAS_GEMM_KERNEL=0: k_gemm, 1: k_gemm_nb): time + bit equality of C.
usage: python tools/exp/kgemm_ab.py [size=2048] [kind=D113]   (spawns one process per kernel: the choice is read once)"""
import os, sys, subprocess, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("KGEMM_CHILD") is None:
    for k in ("0", "1"):
        env = dict(os.environ, KGEMM_CHILD="1", QBLAS_GEMM_KERNEL=k)
        subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, check=False)
    sys.exit(0)
import torch
import qblas_b200 as qb
from gpu_util import dev_random
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
kind = sys.argv[2] if len(sys.argv) > 2 else "D113"
dev = torch.device("cuda:0"); torch.cuda.set_device(0); qb.init()
qb.set_mode(qb.MODE_REFERENCE)
A = dev_random((S * S,), kind, 1, dev); B = dev_random((S * S,), kind, 2, dev); C0 = dev_random((S * S,), kind, 3, dev)
ts = []
for it in range(4):
    C = C0.clone()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); qb.gemm("R", S, S, S, 1.5, A, S, B, S, 0.75, C, S); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
h = hashlib.sha256(C.cpu().numpy().tobytes()).hexdigest()[:16]
ms = min(ts[1:])
print(f"kernel {os.environ['QBLAS_GEMM_KERNEL']} {kind} {S}^3: {ms:.2f} ms = {2.0 * S ** 3 / ms / 1e6:.1f} GFLOP/s  sha {h}", flush=True)
