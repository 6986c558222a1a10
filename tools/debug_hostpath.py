import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import qblas_b200 as qb
import qgen, oracle_lib
from gpu_util import to_dev, to_host
from qblas_b200 import quad
qb.init(); orc = oracle_lib.load_oracle()
m, n, k = 4200, 520, 640
rng = np.random.default_rng(m + n)
lda, ldb, ldc = k + 3, n + 1, n + 2
A = qgen.matrix(rng, m, k, "D113", lda); B = qgen.matrix(rng, k, n, "D113", ldb); C0 = qgen.matrix(rng, m, n, "D113", ldc)
alpha, beta = quad.random_quads(rng, 2)
qb.set_mode(qb.MODE_FAST)
dC = to_dev(C0)
qb.gemm("R", m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc); torch.cuda.synchronize()
print("device plan", qb.oz_last_stats())
dev = to_host(dC)
for trial in range(3):
    hC = C0.copy()
    qb.gemm("R", m, n, k, alpha, A, lda, B, ldb, beta, hC, ldc)
    st = qb.oz_last_stats()
    bad = np.argwhere(~quad.same_bits(dev, hC)).reshape(-1)
    print("host plan", st, "diffs", len(bad))
    for e in bad[:12]:
        i, j = divmod(int(e), ldc)
        idx = np.array([[i, j]])
        exact, _, _ = orc.exact_dot_check("R", k, A, lda, B, ldb, idx)
        al = np.asarray(alpha, dtype=np.uint64).reshape(1, 2); be = np.asarray(beta, dtype=np.uint64).reshape(1, 2)
        want = orc.fma(al, exact, orc.mul(be, np.ascontiguousarray(C0[e:e + 1])))
        print(f"  ({i},{j}) in-row-range={j < n} dev==want {bool(quad.same_bits(dev[e:e+1], want).all())} host==want {bool(quad.same_bits(hC[e:e+1], want).all())} host==C0 {bool((hC[e] == C0[e]).all())}"
              f" dev {dev[e]} host {hC[e]}")
# device path with the check forced on (window below the span): do the same entries move?
qb.set_mode(qb.MODE_REFERENCE)
