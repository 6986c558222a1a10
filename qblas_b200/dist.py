"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md §8e:

  qgemm : C (and A) split into contiguous row blocks; B broadcast as bytes; C blocks all-gathered
          as bytes.  The per-element reduction order does not depend on the row split, so the
          result is bitwise the 1-GPU result in both modes.
  qgemv : row blocks of A / y; x broadcast; y blocks all-gathered.
  qdot  : contiguous index ranges.  NCCL has no binary128 sum, so there is NO all-reduce: the
          16-byte partials are all-gathered as bytes and folded on every rank in a fixed order
          with the library's own add.  In reference-order mode the ranges are whole reference
          chunks (n/T each, level1.hpp:46-53), so the fold over all T partials in tid order is
          exactly the reference's.

`compute` is the per-rank engine (default: the CUDA library through qblas_b200.api).  The CPU
tests inject a stand-in so that the partition / collective logic is covered without a GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def row_block(m: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `m` rows for `rank`; the first m % world ranks get one extra row."""
    base, rem = divmod(m, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_block(T: int, world: int, rank: int):
    """Contiguous block of reference-order dot chunks [c0, c1) owned by `rank`."""
    return row_block(T, world, rank)


def _bytes(t: torch.Tensor) -> torch.Tensor:
    return t.view(torch.uint8)


class PeerBuffer:
    """One equally sized device buffer per rank, each mapped into every other rank's address space (CUDA IPC through the
    library's qb_peer_* calls; all ranks on one NVLink node).  `tensor` is the local buffer as a torch uint8 tensor,
    `ptrs[q]` the address of rank q's buffer as seen from this process (ptrs[rank] = the local one).

    Collective, and failures are collective too: if the allocation, the export or the mapping fails on ANY rank, every
    rank releases what it holds and raises RuntimeError, so that all of them fall back to the NCCL gather together
    (a rank that left alone would leave the others waiting in a collective).

    `mem` is the memory backend (default: qblas_b200.api = the CUDA library); the CPU tests inject a shared-memory
    stand-in with the same five calls (peer_alloc / peer_export / peer_open / peer_close / peer_free)."""

    def __init__(self, nbytes, group=None, mem=None):
        if mem is None:
            from . import api as mem
        self.mem, self.group = mem, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        self.local_ptr, self.tensor, self.ptrs = None, None, []
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        err, handle = None, bytes(64)
        try:
            self.local_ptr, self.tensor = mem.peer_alloc(self.nbytes)
            handle = bytes(mem.peer_export(self.local_ptr))
            assert len(handle) == 64
        except Exception as e:   # reported to every rank below
            err = e
        mine = torch.tensor(list(handle) + [0 if err is None else 1], dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        gathered = [bytes(h.cpu().tolist()) for h in gathered]
        bad = [q for q in range(self.world) if gathered[q][64]]
        if bad:
            self._release([])
            raise RuntimeError(f"peer buffer: allocation / export failed on rank(s) {bad}" + (f": {err!r}" if err is not None else ""))
        opened = []
        try:
            for q in range(self.world):
                if q == self.rank:
                    self.ptrs.append(self.local_ptr)
                else:
                    self.ptrs.append(mem.peer_open(gathered[q][:64]))
                    opened.append(self.ptrs[-1])
        except Exception as e:
            err = e
        flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()):
            self._release(opened)
            raise RuntimeError("peer buffer: mapping a peer's memory failed on some rank" + (f" (here: {err!r})" if err is not None else ""))

    def _release(self, opened):
        for p in opened:
            try:
                self.mem.peer_close(p)
            except Exception:
                pass
        self.ptrs = []
        self.tensor = None
        if self.local_ptr is not None:
            self.mem.peer_free(self.local_ptr)
            self.local_ptr = None

    def close(self):
        if self.tensor is not None and self.tensor.is_cuda:
            torch.cuda.synchronize()
        dist.barrier(group=self.group)          # nobody still writes into a buffer that is about to go
        self._release([p for q, p in enumerate(self.ptrs) if q != self.rank])


class SymmetricBuffer:
    """Same role as PeerBuffer, on torch's symmetric memory (torch.distributed._symmetric_memory: the allocation, the handle exchange
    and the NVSwitch multicast mapping are torch plumbing).  `ptrs[q]` = rank q's buffer as seen from this process, and `mc_ptr` = ONE
    address that the NVSwitch replicates to every rank's buffer (0 when the fabric has no multicast): a kernel that stores a finished
    C element there once reaches all the copies, so the egress of the fused gather is 1x the block instead of (world - 1)x.
    Raises RuntimeError where symmetric memory is unavailable (gloo groups, old drivers): callers fall back to PeerBuffer / NCCL."""

    def __init__(self, nbytes, group=None):
        if dist.get_backend(group) != "nccl":
            raise RuntimeError("symmetric memory needs CUDA devices (nccl group)")
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.group = group if group is not None else dist.group.WORLD
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self.nbytes = int(nbytes)
            self.tensor = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
            self.handle = symm_mem.rendezvous(self.tensor, self.group)
            self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
            self.mc_ptr = int(self.handle.multicast_ptr or 0)
            self.local_ptr = self.ptrs[self.rank]
        except Exception as e:
            raise RuntimeError(f"symmetric memory unavailable: {e!r}")

    def close(self):
        if self.tensor is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        self.tensor = None
        self.handle = None


def fused_outputs(buf, rank, byte_offset, multicast=True):
    """Addresses for qb_set_gemm_peer_outputs: the multicast address when the buffer has one (one store reaches every rank), else the
    other ranks' mapped buffers."""
    if multicast and getattr(buf, "mc_ptr", 0):
        return [buf.mc_ptr + byte_offset]
    return [buf.ptrs[q] + byte_offset for q in range(len(buf.ptrs)) if q != rank]


class SharedPlanes:
    """Residue planes of B shared between the ranks of a row-sharded qgemm (fast mode).  Every rank of such a product needs the int8
    residue planes of ALL of B; instead of every rank reducing the whole of B (the one part of the work that does not shrink with the
    number of GPUs), rank r reduces one column slice of each panel and stores it ONCE to the NVSwitch multicast address of a symmetric
    [N][n][Kp] buffer, which puts it into every rank's copy; a symmetric-memory barrier per panel tells the ranks that the panel is
    complete.  Needs torch symmetric memory with multicast; holds its buffers across calls (grow-only)."""

    def __init__(self, group=None):
        self.group, self.buf, self.stage, self.stream, self.key = group, None, None, None, None

    def ensure(self, N, n, Kp, slice_cols):
        key = (N, n, Kp)
        if self.buf is None or self.key != key:
            if self.buf is not None:
                self.buf.close()
            self.buf = SymmetricBuffer(N * n * Kp, group=self.group)
            if not self.buf.mc_ptr:
                raise RuntimeError("no multicast address")
            self.key = key
        need = N * slice_cols * Kp
        if self.stage is None or self.stage.numel() < need:
            self.stage = torch.empty(need, dtype=torch.int8, device=self.buf.tensor.device)
        if self.stream is None:
            self.stream = torch.cuda.Stream(priority=0)
        return self.buf

    def close(self):
        if self.buf is not None:
            self.buf.close()
        self.buf = self.stage = None


def _lines_span(stats, count):
    """widest bit span of the `count` lines described by a statistics block [emax | lmin | sp]; any Inf/NaN"""
    emax, lmin, sp = stats[:count], stats[count:2 * count], stats[2 * count:3 * count]
    span = torch.where(emax != 0, emax + 113 - lmin, torch.zeros_like(emax)).max()
    return span, sp.max()


class _CudaEngine:
    """Default per-rank engine: libqblas_b200.so on the current CUDA device."""

    def gemm(self, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, on_rows=None, min_passes=1):
        """on_rows(row0, rows): called after the work producing C rows [row0, row0 + rows) has been enqueued on the current stream
        (every row exactly once); with min_passes > 1 the tensor path produces the rows in that many passes."""
        from . import api
        if on_rows is None:
            api.gemm("R", m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
            return
        api.set_gemm_pass_callback(on_rows, min_passes)
        try:
            api.gemm("R", m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        finally:
            api.set_gemm_pass_callback(None)

    def gemm_streamed(self, m, n, k, alpha, A, lda, panel_fn, panel_cols, colstats, beta, C, ldc, planes=None):
        """qgemm whose B arrives in column panels: panel_fn(col0, cols, stream_ptr) -> (ptr, ld) (qb_set_gemm_b_panels).
        planes = (moduli, window): the panels are residue planes (qb_set_gemm_b_planes)."""
        from . import api
        api.set_gemm_b_panels(panel_fn, panel_cols, colstats)
        if planes:
            api.set_gemm_b_planes(*planes)
        try:
            api.gemm("R", m, n, k, alpha, A, lda, A, n, beta, C, ldc)      # the B argument is not dereferenced
        finally:
            api.set_gemm_b_planes(0)
            api.set_gemm_b_panels(None)

    def colstats(self, k, n, B, ldb, out):
        from . import api
        return api.gemm_colstats("R", k, n, B, ldb, out)

    def wait_on_stream(self, work, stream_ptr):
        """make the library's CUDA stream wait for an async collective"""
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            work.wait()

    def set_peer_outputs(self, ptrs):
        """fused gather: addresses of the C block inside the peers' buffers for the following gemm calls (None: off)"""
        from . import api
        api.set_gemm_peer_outputs(ptrs)

    def peer_written(self):
        """peers the last gemm stored to (0: it ran a path without the fused stores)"""
        from . import api
        return api.gemm_peer_written()

    def gemv(self, m, n, alpha, A, lda, x, beta, y, m_total=0):
        from . import api
        api.gemv("R", m, n, alpha, A, lda, x, 1, beta, y, 1, m_total=m_total)

    def dot_partials(self, n_local, x, y, chunk, nchunks, out):
        from . import api
        api.dot_partials(n_local, x, 1, y, 1, chunk, nchunks, out)

    def dot_fast(self, n_local, x, y, out):
        from . import api
        api.dot(n_local, x, 1, y, 1, out)

    def fold(self, count, partials, out, do_sqrt=False):
        from . import api
        api.fold_partials(count, partials, out, do_sqrt)


def _gather_blocks(m, n, world, C_full, C_blk, group):
    if m % world == 0:
        dist.all_gather_into_tensor(_bytes(C_full), _bytes(C_blk), group=group)
    else:  # ragged: one broadcast per owner (grouped broadcasts)
        for r in range(world):
            l2, h2 = row_block(m, world, r)
            if h2 > l2:
                dist.broadcast(_bytes(C_full[l2 * n:h2 * n]), src=r, group=group)


def qgemm_row_sharded(m, n, k, alpha, A_blk, B, beta, C_full, *, src=0, compute=None, group=None, overlap_passes=1, peers=None, b_panels=0,
                      b_packed=None, share_planes=None):
    """C_full (m x n, row-major, identical buffer shape on every rank) <- alpha*A*B + beta*C.
    A_blk holds this rank's rows [lo, hi) of A (row-major, lda = k); B (k x n) is valid on `src`
    and is overwritten by the broadcast elsewhere.  Returns (lo, hi).

    overlap_passes > 1 (even split only): the local block is produced in that many row passes and the all-gather of each
    pass's rows is issued from the library's row-pass hook as soon as that pass is enqueued, so it runs on NCCL's stream
    while the next pass computes; only the last pass's gather is exposed.  Same bytes, same result.

    peers = a PeerBuffer / SymmetricBuffer whose local tensor IS C_full's storage (fused gather): the kernel that finishes the C elements
    stores them into every rank's C_full over NVLink as well (through the NVSwitch multicast address when the buffer has one), no
    all-gather is issued, and a barrier makes the step complete.  If the library ran a path without the fused stores (it reports 0
    outputs written) the NCCL all-gather is issued after all.

    b_panels = w > 0 (fast mode, every rank owns >= 1 row): B is not broadcast before the product but DURING it, in column panels of w
    columns (a multiple of 256): the owner computes the column statistics of B (12 n bytes, broadcast first), packs panel j into a
    contiguous k x w buffer and broadcasts it while the ranks multiply panel j-1 (qb_set_gemm_b_panels).  b_packed: optional
    preallocated (k * n, 2) buffer for the packed panels (panel j occupies rows [j*k*w, ...)); B itself is only read on `src`.

    share_planes = a SharedPlanes object (with b_panels, NCCL, multicast fabric): the ranks do not each reduce all of B to residue
    planes; rank r reduces one column slice of every panel and multicasts the planes (see SharedPlanes).  Used only when the windows
    cover the spans on every rank (exact case: nothing for the fix-up kernel, which would need B's elements); otherwise, and whenever
    the set-up fails on any rank, the call proceeds with element panels."""
    compute = compute or _CudaEngine()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_block(m, world, rank)
    C_blk = C_full[lo * n:hi * n]
    streamed = b_panels > 0 and world > 1 and m >= world
    panel_fn = stats = None
    if streamed:
        w = int(b_panels)
        dev, dt = C_full.device, B.dtype
        stats = torch.zeros(3 * n, dtype=torch.int32, device=dev)
        if rank == src:
            compute.colstats(k, n, B, n, stats)
        dist.broadcast(stats, src=src, group=group)
        packed = b_packed if b_packed is not None else torch.empty((k * n, 2), dtype=dt, device=dev)
        B3 = B.reshape(k, n, 2) if rank == src else None
        works, views = [], []
        for c0 in range(0, n, w):
            cols = min(w, n - c0)
            pv = packed[c0 * k:(c0 + cols) * k]                      # panel j: k x cols, contiguous, ld = cols
            if rank == src:
                pv.reshape(k, cols, 2).copy_(B3[:, c0:c0 + cols])
            works.append(dist.broadcast(_bytes(pv), src=src, group=group, async_op=True))
            views.append(pv)

        def panel_fn(col0, cols, stream_ptr):
            j = col0 // w
            compute.wait_on_stream(works[j], stream_ptr)
            return views[j].data_ptr(), cols

        if share_planes is not None and isinstance(compute, _CudaEngine) and hi > lo:
            planes_mode = _share_planes_setup(share_planes, m, n, k, w, lo, hi, A_blk, stats, works, views, world, rank, group)
            if planes_mode is not None:
                panel_fn, planes_mode = planes_mode
    else:
        dist.broadcast(_bytes(B), src=src, group=group)
    if not streamed or share_planes is None or not isinstance(compute, _CudaEngine) or hi <= lo:
        planes_mode = None

    def run_gemm(rows, A_, C_, **kw):
        if streamed and planes_mode is not None:
            compute.gemm_streamed(rows, n, k, alpha, A_, k, panel_fn, int(b_panels), stats, beta, C_, n, planes=planes_mode)
        elif streamed:
            compute.gemm_streamed(rows, n, k, alpha, A_, k, panel_fn, int(b_panels), stats, beta, C_, n)
        else:
            compute.gemm(rows, n, k, alpha, A_, k, B, n, beta, C_, n, **kw)

    if peers is not None and world > 1:
        expect = -1                                                  # a rank without rows has nothing to deliver
        wrote = -1
        if hi > lo:
            outs = fused_outputs(peers, rank, lo * n * 16)
            expect = len(outs)
            compute.set_peer_outputs(outs)
            try:
                run_gemm(hi - lo, A_blk, C_blk)
                wrote = compute.peer_written()
            finally:
                compute.set_peer_outputs(None)
        # every rank must take the same branch: a rank whose planner declined makes all of them gather
        flag = torch.tensor([1 if wrote == expect else 0], dtype=torch.int32, device=C_full.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also the completion barrier of the peer stores
        if int(flag.item()) == 1:
            return lo, hi
        _gather_blocks(m, n, world, C_full, C_blk, group)            # the blocks are computed; only the exchange is left
        return lo, hi
    if m % world == 0 and overlap_passes > 1 and hi > lo and not streamed:
        m_loc = hi - lo
        works2 = []

        def on_rows(r0, rows):
            # rows [r0, r0 + rows) of EVERY rank's block: rank q's copy lands at rows q * m_loc + r0 of C_full
            outs = [_bytes(C_full[(q * m_loc + r0) * n:(q * m_loc + r0 + rows) * n]) for q in range(world)]
            works2.append(dist.all_gather(outs, _bytes(C_blk[r0 * n:(r0 + rows) * n]), group=group, async_op=True))

        compute.gemm(m_loc, n, k, alpha, A_blk, k, B, n, beta, C_blk, n, on_rows=on_rows, min_passes=overlap_passes)
        for wk in works2:
            wk.wait()
        return lo, hi
    if hi > lo:
        run_gemm(hi - lo, A_blk, C_blk)
    elif streamed:
        for wk in works:                                             # a rank without rows still takes part in the broadcasts
            wk.wait()
    _gather_blocks(m, n, world, C_full, C_blk, group)
    return lo, hi


def _share_planes_setup(sp, m, n, k, w, lo, hi, A_blk, stats, works, views, world, rank, group):
    """Shared residue planes of B (see SharedPlanes): returns (panel_fn, (moduli, window_B)) or None when the ranks must fall back to
    element panels (every rank takes the same branch: the decision is all-reduced)."""
    from . import api
    dev = stats.device
    ok, N, WA, WB = 1, 0, 0, 0
    try:
        statsA = torch.zeros(3 * (hi - lo), dtype=torch.int32, device=dev)
        api.gemm_colstats("C", k, hi - lo, A_blk, k, statsA)          # rows of the row-major A block = columns of its col-major view
        spanA, flA = _lines_span(statsA, hi - lo)
        spanB, flB = _lines_span(stats, n)
        red = torch.stack([spanA, spanB, flA, flB]).to(torch.int64)
    except Exception:
        ok, red = 0, torch.zeros(4, dtype=torch.int64, device=dev)
    dist.all_reduce(red, op=dist.ReduceOp.MAX, group=group)            # every rank plans with the widest A rows of all ranks
    sa, sb_, fa, fb = (int(v) for v in red.tolist())
    plan = api.crt_plan(sa, sb_, k) if ok else None
    if plan is None or plan[3] != 0 or fa or fb:
        ok = 0
    else:
        N, WA, WB, _ = plan
    Kp = (k + 127) // 128 * 128
    sw = -(-w // world)
    if ok:
        try:
            buf = sp.ensure(N, n, Kp, sw)
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        return None
    mc = api.tensor_from_ptr(buf.mc_ptr, N * n * Kp).view(torch.int64).reshape(N, n, Kp // 8)
    emaxB = stats[:n]
    events = []
    cur = torch.cuda.current_stream()
    sp.stream.wait_stream(cur)
    with torch.cuda.stream(sp.stream):
        for j, c0 in enumerate(range(0, n, w)):
            cols = min(w, n - c0)
            s0 = min(cols, rank * sw); cnt = max(0, min(cols, (rank + 1) * sw) - s0)
            works[j].wait()                                            # the packed k x cols panel has arrived (side stream waits)
            if cnt > 0:
                st = sp.stage[:N * cnt * Kp]
                # my slice of the panel: columns s0 .. s0 + cnt of the packed panel (ld = cols)
                api.crt_residues("R", k, cnt, None, cols, emaxB[c0 + s0:], WB, N, st, 0, data_ptr=views[j].data_ptr() + s0 * 16)
                mc[:, c0 + s0:c0 + s0 + cnt].copy_(st.view(torch.int64).reshape(N, cnt, Kp // 8))
            buf.handle.barrier(channel=0)                              # every rank's slice of this panel has landed everywhere
            ev = torch.cuda.Event(); ev.record(sp.stream)
            events.append(ev)
    base = buf.local_ptr

    def panel_fn(col0, cols, stream_ptr):
        torch.cuda.ExternalStream(stream_ptr).wait_event(events[col0 // w])
        return base + col0 * Kp, n * Kp

    return panel_fn, (N, WB)


def qgemv_row_sharded(m, n, alpha, A_blk, x, beta, y_full, *, src=0, compute=None, group=None):
    compute = compute or _CudaEngine()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_block(m, world, rank)
    dist.broadcast(_bytes(x), src=src, group=group)
    if hi > lo:
        compute.gemv(hi - lo, n, alpha, A_blk, n, x, beta, y_full[lo:hi], m_total=m)   # planned as the whole qgemv: same bits as one GPU
    for r in range(world):
        l2, h2 = row_block(m, world, r)
        if h2 > l2:
            dist.broadcast(_bytes(y_full[l2:h2]), src=r, group=group)
    return lo, hi


def dot_shard_range(n: int, T: int, world: int, rank: int, reference_order: bool):
    """Element range [lo, hi) of the vectors that `rank` must hold."""
    if reference_order and n >= 500:
        chunk = n // T
        c0, c1 = chunk_block(T, world, rank)
        lo = c0 * chunk
        hi = n if c1 == T else c1 * chunk
        return lo, hi
    if reference_order:          # n < 500: the reference runs one kernel over everything (level1.hpp:40)
        return (0, n) if rank == 0 else (0, 0)
    return row_block(n, world, rank)


def qdot_sharded(n, x_loc, y_loc, T, out, *, reference_order=True, do_sqrt=False, compute=None, group=None):
    """x_loc / y_loc: this rank's slice (see dot_shard_range).  `out`: 16-byte device buffer that
    receives the full result on EVERY rank (all ranks fold the same gathered partials)."""
    compute = compute or _CudaEngine()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev, dt = x_loc.device, x_loc.dtype
    if reference_order:
        if n < 500:
            Tn, chunk = 1, n
        else:
            Tn, chunk = T, n // T
        c0, c1 = chunk_block(Tn, world, rank)
        per = -(-Tn // world)  # padded slots per rank so that one all_gather serves ragged splits
        mine = torch.zeros((per, 2), dtype=dt, device=dev)
        lo, hi = dot_shard_range(n, T, world, rank, True)
        if c1 > c0:
            compute.dot_partials(hi - lo, x_loc, y_loc, chunk, c1 - c0, mine[:c1 - c0])
        gathered = torch.zeros((world * per, 2), dtype=dt, device=dev)
        dist.all_gather_into_tensor(_bytes(gathered), _bytes(mine), group=group)
        # compact to tid order (drop the padding slots of ragged splits)
        parts = torch.cat([gathered[r * per:r * per + (chunk_block(Tn, world, r)[1] - chunk_block(Tn, world, r)[0])] for r in range(world)])
        if n < 500:
            # the reference returns the single kernel result directly (no +0 fold)
            out.copy_(parts[:1].reshape(out.shape)) if not do_sqrt else compute.fold(1, parts, out, True)
        else:
            compute.fold(Tn, parts.contiguous(), out, do_sqrt)
        return out
    mine = torch.zeros((1, 2), dtype=dt, device=dev)
    lo, hi = dot_shard_range(n, T, world, rank, False)
    compute.dot_fast(hi - lo, x_loc, y_loc, mine)
    gathered = torch.zeros((world, 2), dtype=dt, device=dev)
    dist.all_gather_into_tensor(_bytes(gathered), _bytes(mine), group=group)
    compute.fold(world, gathered, out, do_sqrt)
    return out
