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

    def __init__(self, mem=None, fused_ok=True):
        import oracle_lib
        self.o = oracle_lib.load_oracle()
        self.mem, self.fused_ok, self.peers, self.wrote = mem, fused_ok, [], 0

    # fused gather (qb_set_gemm_peer_outputs / qb_get_gemm_peer_written): the stand-in copies the finished block into the peers' memory
    def set_peer_outputs(self, ptrs):
        self.peers = list(ptrs or [])

    def peer_written(self):
        return self.wrote

    @staticmethod
    def _np(t):
        return t.numpy().view(np.uint64)

    def gemm(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, on_rows=None, min_passes=1):
        self.o.gemm("R", m, n, k, alpha, self._np(A), lda, self._np(B), ldb, beta, self._np(C), ldc)
        self.wrote = 0
        if self.peers and self.fused_ok:
            for p in self.peers:
                self.mem.view(p, m * n * 16).copy_(C.view(torch.uint8).reshape(-1)[:m * n * 16])
            self.wrote = len(self.peers)
        if on_rows is not None:   # the library's row-pass hook: every row reported exactly once, in passes
            step = -(-m // max(1, min_passes))
            for r0 in range(0, m, step):
                on_rows(r0, min(step, m - r0))

    # streamed B (qb_set_gemm_b_panels): the stand-in asks for every panel in order, as the library does, and multiplies panel by panel
    def colstats(self, k, n, B, ldb, out):
        out.fill_(7)            # opaque to the host logic: it is only broadcast and handed back
        return out

    def wait_on_stream(self, work, stream_ptr):
        work.wait()

    def gemm_streamed(self, m, n, k, alpha, A, lda, panel_fn, panel_cols, colstats, beta, C, ldc):
        import ctypes
        assert (colstats.numpy() == 7).all() and colstats.numel() == 3 * n       # the owner's statistics arrived on every rank
        Cn = self._np(C)
        for c0 in range(0, n, panel_cols):
            cols = min(panel_cols, n - c0)
            ptr, ld = panel_fn(c0, cols, 0)
            assert ld == cols
            Bp = np.ctypeslib.as_array((ctypes.c_uint64 * (k * cols * 2)).from_address(ptr)).reshape(k * cols, 2)
            Cp = np.ascontiguousarray(Cn.reshape(m, ldc, 2)[:, c0:c0 + cols]).reshape(m * cols, 2)
            self.o.gemm("R", m, cols, k, alpha, self._np(A), lda, np.ascontiguousarray(Bp), cols, beta, Cp, cols)
            Cn.reshape(m, ldc, 2)[:, c0:c0 + cols] = Cp.reshape(m, cols, 2)
        self.wrote = 0
        if self.peers and self.fused_ok:
            for p in self.peers:
                self.mem.view(p, m * n * 16).copy_(C.view(torch.uint8).reshape(-1)[:m * n * 16])
            self.wrote = len(self.peers)

    def gemv(self, m, n, alpha, A, lda, x, beta, y, m_total=0):
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


class ShmMem:
    """CPU stand-in for the library's qb_peer_* calls: the 'device memory' is a file under /dev/shm mapped MAP_SHARED by every rank,
    the 64-byte handle is its path, a 'pointer' is (owner rank + 1) << 40 plus a byte offset."""

    def __init__(self, tag, rank, fail_alloc=False, fail_open=False):
        self.tag, self.rank, self.fail_alloc, self.fail_open = tag, rank, fail_alloc, fail_open
        self.maps, self.paths = {}, {}

    def peer_alloc(self, nbytes):
        if self.fail_alloc:
            raise MemoryError("injected allocation failure")
        path = f"/dev/shm/qbtest_{self.tag}_{self.rank}"
        with open(path, "wb") as f:
            f.truncate(nbytes)
        t = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
        slot = self.rank + 1
        self.maps[slot], self.paths[slot] = t, path
        return slot << 40, t

    def peer_export(self, ptr):
        return self.paths[ptr >> 40].encode().ljust(64, b"\0")

    def peer_open(self, handle):
        if self.fail_open:
            raise OSError("injected mapping failure")
        path = bytes(handle).rstrip(b"\0").decode()
        t = torch.from_file(path, shared=True, size=os.path.getsize(path), dtype=torch.uint8)
        slot = int(path.rsplit("_", 1)[1]) + 1
        self.maps[slot] = t
        return slot << 40

    def peer_close(self, ptr):
        self.maps.pop(ptr >> 40, None)

    def peer_free(self, ptr):
        self.maps.pop(ptr >> 40, None)
        path = self.paths.pop(ptr >> 40, None)
        if path and os.path.exists(path):
            os.unlink(path)

    def view(self, ptr, nbytes):
        off = ptr & ((1 << 40) - 1)
        return self.maps[ptr >> 40][off:off + nbytes]


def _fused_worker(rank, world, port, q):
    """dist.PeerBuffer + qgemm_row_sharded(peers=...) on CPU: shared-memory peers, the oracle as the engine."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qblas_b200 import dist as qd, quad
        import oracle_lib
        orc = oracle_lib.load_oracle()
        rng = np.random.default_rng(17)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64))
        tag = f"{port}"
        for (m, n, k), fused_ok_on in (((8, 5, 40), None), ((7, 6, 20), None), ((8, 5, 40), 1), ((2, 3, 9), None)):
            A = quad.random_quads(rng, m * k); B = quad.random_quads(rng, k * n); C0 = quad.random_quads(rng, m * n)
            alpha, beta = quad.random_quads(rng, 2)
            want = C0.copy(); orc.gemm("R", m, n, k, alpha, A, k, B, n, beta, want, n)
            mem = ShmMem(tag, rank)
            # fused_ok_on = r: rank r's engine reports 0 peers written (a declined planner) -> every rank must gather instead
            eng = OracleEngine(mem=mem, fused_ok=(fused_ok_on is None or fused_ok_on != rank))
            pb = qd.PeerBuffer(m * n * 16, mem=mem)
            assert len(pb.ptrs) == world and pb.ptrs[rank] == pb.local_ptr
            Cp = pb.tensor.view(torch.int64).reshape(m * n, 2)
            Cp.copy_(t(C0))
            dist.barrier()
            lo, hi = qd.row_block(m, world, rank)
            Bt = t(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64)
            qd.qgemm_row_sharded(m, n, k, alpha, t(A[lo * k:hi * k]), Bt, beta, Cp, compute=eng, peers=pb, b_panels=(3 if m >= world and k == 40 else 0))
            dist.barrier()
            assert (Cp.numpy().view(np.uint64) == want).all(), ("fused gemm", m, n, k, rank, fused_ok_on)
            assert eng.peers == []                                   # switched off again after the call
            del Cp
            pb.close()
            assert not os.path.exists(f"/dev/shm/qbtest_{tag}_{rank}")
        # failures are collective: whichever rank fails, EVERY rank raises (and nobody is left waiting in a collective)
        for kw in ({"fail_alloc": rank == world - 1}, {"fail_open": rank == 0}):
            mem = ShmMem(tag, rank, **kw)
            try:
                qd.PeerBuffer(64, mem=mem)
                raise AssertionError(("PeerBuffer did not raise", kw, rank))
            except RuntimeError:
                pass
            dist.barrier()
            assert not os.path.exists(f"/dev/shm/qbtest_{tag}_{rank}"), "local buffer leaked after a failed setup"
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


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
            # B broadcast DURING the product, in packed column panels (here 2 columns wide, ragged last panel), statistics first
            Bt = t(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64)
            Cf = t(C0.copy())
            qd.qgemm_row_sharded(m, n, k, alpha, t(A[lo * k:hi * k]), Bt, beta, Cf, compute=eng, b_panels=2)
            assert (Cf.numpy().view(np.uint64) == want).all(), ("gemm streamed B", m, n, k, rank)
            if rank != 0:
                assert (Bt == 0).all()                                   # B itself is only read on its owner
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


@pytest.mark.parametrize("world", [2, 3])
def test_fused_gather_plumbing_and_collective_failures(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + world + os.getpid() % 200
    procs = [ctx.Process(target=_fused_worker, args=(r, world, port, q)) for r in range(world)]
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
