"""N > 1 host logic on CPU: world_size-2 (and 3, ragged) gloo runs of qblas_b200.dist with the ORACLE as the
per-rank engine (test infrastructure standing in for the CUDA library, which needs a GPU).  Checks the
partitioning / broadcast / all-gather / fixed-order fold plumbing: sharded results must equal the
single-process oracle bit for bit (row sharding does not change any element's order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """stand-in engine: same call surface as qblas_b200.dist._CudaEngine, computed by oracle/qoracle.c"""

    def __init__(self):
        import oracle_lib
        self.o = oracle_lib.load_oracle()

    @staticmethod
    def _np(t):
        return t.numpy().view(np.uint64)

    def gemm(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, on_rows=None, min_passes=1):
        self.o.gemm("R", m, n, k, alpha, self._np(A), lda, self._np(B), ldb, beta, self._np(C), ldc)
        if on_rows is not None:   # the library's row-pass hook: every row reported exactly once, in passes
            step = -(-m // max(1, min_passes))
            for r0 in range(0, m, step):
                on_rows(r0, min(step, m - r0))

    def gemv(self, m, n, alpha, A, lda, x, beta, y):
        self.o.gemv("R", m, n, alpha, self._np(A), lda, self._np(x), 1, beta, self._np(y), 1)

    def dot_partials(self, n_local, x, y, chunk, nchunks, out):
        xs, ys = self._np(x), self._np(y)
        for c in range(nchunks):
            s = c * chunk
            e = n_local if c == nchunks - 1 else s + chunk
            # one reference chunk = the two-lane kernel = the oracle's dot with T = 1 on n < 500 semantics
            out[c] = torch.from_numpy(self.o.dot(e - s, xs[s:e], 1, ys[s:e], 1, 1).view(np.int64)) if e - s < 500 else \
                torch.from_numpy(self._chunk(xs[s:e], ys[s:e]).view(np.int64))

    def _chunk(self, xs, ys):
        # n >= 500 with T = 1: one chunk, then add(+0, partial): identical bits unless partial == -0 (cannot happen from a +0 seed)
        return self.o.dot(xs.shape[0], xs, 1, ys, 1, 1)

    def dot_fast(self, n_local, x, y, out):
        out[0] = torch.from_numpy(self.o.dot(n_local, self._np(x), 1, self._np(y), 1, 1).view(np.int64))

    def fold(self, count, partials, out, do_sqrt=False):
        p = self._np(partials)
        r = np.zeros(2, dtype=np.uint64)
        for t in range(count):
            r = self.o.add(r.reshape(1, 2), np.ascontiguousarray(p[t]).reshape(1, 2))[0]
        if do_sqrt:
            r = self.o.sqrt(r.reshape(1, 2))[0]
        out.copy_(torch.from_numpy(r.view(np.int64)).reshape(out.shape))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qblas_b200 import dist as qd, quad
        import oracle_lib
        orc = oracle_lib.load_oracle()
        eng = OracleEngine()
        rng = np.random.default_rng(11)                       # same stream on every rank
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64))
        # ---- qgemm, even and ragged row split
        for m, n, k in ((8, 5, 130), (7, 6, 20)):
            A = quad.random_quads(rng, m * k); B = quad.random_quads(rng, k * n); C0 = quad.random_quads(rng, m * n)
            alpha, beta = quad.random_quads(rng, 2)
            want = C0.copy(); orc.gemm("R", m, n, k, alpha, A, k, B, n, beta, want, n)
            lo, hi = qd.row_block(m, world, rank)
            Bt = t(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64)
            Cf = t(C0.copy())
            qd.qgemm_row_sharded(m, n, k, alpha, t(A[lo * k:hi * k]), Bt, beta, Cf, compute=eng)
            assert (Cf.numpy().view(np.uint64) == want).all(), ("gemm", m, n, k, rank)
            # same call with the all-gather issued pass by pass from the row-pass hook (even splits only; ragged falls back)
            Bt = t(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64)
            Cf = t(C0.copy())
            qd.qgemm_row_sharded(m, n, k, alpha, t(A[lo * k:hi * k]), Bt, beta, Cf, compute=eng, overlap_passes=2)
            assert (Cf.numpy().view(np.uint64) == want).all(), ("gemm overlapped", m, n, k, rank)
        # ---- qgemv
        m, n = 9, 33
        A = quad.random_quads(rng, m * n); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
        want = y0.copy(); orc.gemv("R", m, n, 1.5, A, n, x, 1, 0.5, want, 1)
        lo, hi = qd.row_block(m, world, rank)
        xt = t(x) if rank == 0 else torch.zeros((n, 2), dtype=torch.int64)
        yf = t(y0.copy())
        qd.qgemv_row_sharded(m, n, 1.5, t(A[lo * n:hi * n]), xt, 0.5, yf, compute=eng)
        assert (yf.numpy().view(np.uint64) == want).all(), ("gemv", rank)
        # ---- qdot, reference order: T chunks mapped to ranks, folded in tid order on every rank
        for nn, T in ((1200, 4), (1301, 5), (499, 4)):
            x = quad.random_quads(rng, nn); y = quad.random_quads(rng, nn)
            want = orc.dot(nn, x, 1, y, 1, T)
            lo, hi = qd.dot_shard_range(nn, T, world, rank, True)
            out = torch.zeros((1, 2), dtype=torch.int64)
            qd.qdot_sharded(nn, t(x[lo:hi]), t(y[lo:hi]), T, out, reference_order=True, compute=eng)
            assert (out.numpy().view(np.uint64).reshape(2) == want).all(), ("dot", nn, T, rank)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_routines_match_single_process_bitwise(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_row_and_chunk_blocks_partition():
    from qblas_b200 import dist as qd
    for m in (0, 1, 7, 8, 4096, 32768):
        for w in (1, 2, 3, 4, 8):
            blocks = [qd.row_block(m, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
