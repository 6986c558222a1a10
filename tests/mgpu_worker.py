"""TEST worker (launched by tests/test_gpu_multi.py under torch.distributed.run, one rank per GPU, NCCL):
row-sharded qgemm / qgemv and chunk-sharded qdot through qblas_b200.dist on real GPUs.  The sharded
result must be bitwise the single-GPU result (both modes) and, in reference-order mode, the oracle's."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import qblas_b200 as qb
    from qblas_b200 import dist as qd, quad
    import oracle_lib
    from gpu_util import to_dev, to_host
    qb.init()
    orc = oracle_lib.load_oracle()
    rng = np.random.default_rng(3)
    fails = []
    # ---- qgemm: reference order (bit exact vs oracle) and fast mode (tensor path; bitwise equal to the 1-GPU call)
    for mode, (m, n, k) in ((qb.MODE_REFERENCE, (70, 33, 260)), (qb.MODE_REFERENCE, (37, 20, 127)), (qb.MODE_FAST, (1024, 640, 640)), (qb.MODE_FAST, (1031, 250, 300))):   # fast: every rank's block keeps >= 128 rows up to world 8 (tensor path)
        A = quad.random_quads(rng, m * k); B = quad.random_quads(rng, k * n); C0 = quad.random_quads(rng, m * n)
        alpha, beta = quad.random_quads(rng, 2)
        qb.set_mode(mode)
        one = to_dev(C0.copy())
        qb.gemm("R", m, n, k, alpha, to_dev(A), k, to_dev(B), n, beta, one, n)      # single-GPU result on this rank
        lo, hi = qd.row_block(m, world, rank)
        Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
        Cf = to_dev(C0.copy())
        qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cf)
        torch.cuda.synchronize()
        if not (to_host(Cf) == to_host(one)).all():
            fails.append(("gemm vs 1-GPU", mode, m, n, k))
        if m % world == 0:   # all-gather issued pass by pass from the library's row-pass hook: same bits
            Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
            Cf = to_dev(C0.copy())
            qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cf, overlap_passes=3)
            torch.cuda.synchronize()
            if not (to_host(Cf) == to_host(one)).all():
                fails.append(("gemm overlapped vs 1-GPU", mode, m, n, k))
        # fused gather: C lives in peer-mapped memory and the kernel that finishes the elements stores them into every rank's copy
        # (fast-mode residue scheme); in reference order the library reports 0 peers written and the NCCL gather runs after all
        pb = qd.PeerBuffer(m * n * 16)
        Cp = pb.tensor.view(torch.int64).reshape(m * n, 2)
        Cp.copy_(to_dev(C0.copy()))
        Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize(); dist.barrier()
        qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cp, peers=pb)
        torch.cuda.synchronize()
        if not (to_host(Cp) == to_host(one)).all():
            fails.append(("gemm fused gather vs 1-GPU", mode, m, n, k, qb.gemm_peer_written()))
        if mode == qb.MODE_FAST and hi - lo >= 128 and qb.gemm_peer_written() != world - 1:
            fails.append(("fused gather did not run", mode, m, n, k, qb.gemm_peer_written()))
        del Cp
        pb.close()
        # the same through torch symmetric memory: ONE store per element to the NVSwitch multicast address reaches every rank's copy
        try:
            sb = qd.SymmetricBuffer(m * n * 16)
        except RuntimeError as e:
            sb = None
            if rank == 0:
                print(f"mgpu_worker: symmetric memory unavailable, multicast gather not tested: {e}", flush=True)
        if sb is not None:
            Cs = sb.tensor.view(torch.int64).reshape(m * n, 2)
            Cs.copy_(to_dev(C0.copy()))
            Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
            torch.cuda.synchronize(); dist.barrier()
            qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cs, peers=sb)
            torch.cuda.synchronize()
            if not (to_host(Cs) == to_host(one)).all():
                fails.append(("gemm multicast gather vs 1-GPU", mode, m, n, k, sb.mc_ptr != 0, qb.gemm_peer_written()))
            if rank == 0 and mode == qb.MODE_FAST:
                print(f"mgpu_worker: symmetric buffer multicast={'yes' if sb.mc_ptr else 'no'} outputs written={qb.gemm_peer_written()}", flush=True)
            del Cs
            sb.close()
        # B broadcast during the product in packed column panels (fast mode, tensor path)
        if mode == qb.MODE_FAST and n >= 256:
            Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
            Cf2 = to_dev(C0.copy())
            qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cf2, b_panels=256)
            torch.cuda.synchronize()
            if not (to_host(Cf2) == to_host(one)).all():
                fails.append(("gemm streamed B vs 1-GPU", mode, m, n, k))
        # ... and with the residue planes of B computed cooperatively (one column slice per rank) and exchanged over the multicast fabric
        if mode == qb.MODE_FAST and n >= 256:
            shp = qd.SharedPlanes()
            for rep in range(2):                                     # twice: the second call reuses the symmetric buffers
                Bt = to_dev(B) if rank == 0 else torch.zeros((k * n, 2), dtype=torch.int64, device="cuda")
                Cf3 = to_dev(C0.copy())
                qd.qgemm_row_sharded(m, n, k, alpha, to_dev(A[lo * k:hi * k]), Bt, beta, Cf3, b_panels=256, share_planes=shp)
                torch.cuda.synchronize()
                if not (to_host(Cf3) == to_host(one)).all():
                    fails.append(("gemm shared residue planes vs 1-GPU", mode, m, n, k, rep))
            if rank == 0:
                print(f"mgpu_worker: shared residue planes {'used' if shp.buf is not None else 'NOT used (fallback)'}", flush=True)
            shp.close()
        if mode == qb.MODE_REFERENCE:
            want = C0.copy(); orc.gemm("R", m, n, k, alpha, A, k, B, n, beta, want, n)
            if not quad.same_bits(to_host(Cf), want).all():
                fails.append(("gemm vs oracle", m, n, k))
    qb.set_mode(qb.MODE_REFERENCE)
    # ---- qgemv
    m, n = 1031, 517
    A = quad.random_quads(rng, m * n); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    want = y0.copy(); orc.gemv("R", m, n, 1.5, A, n, x, 1, 0.5, want, 1)
    lo, hi = qd.row_block(m, world, rank)
    xt = to_dev(x) if rank == 0 else torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    yf = to_dev(y0.copy())
    qd.qgemv_row_sharded(m, n, 1.5, to_dev(A[lo * n:hi * n]), xt, 0.5, yf)
    torch.cuda.synchronize()
    if not quad.same_bits(to_host(yf), want).all():
        fails.append(("gemv",))
    # ---- qdot in reference order: T chunks over the ranks, all-gather of 16-byte partials, fixed-order fold
    for nn, T in ((100003, 8), (5000, 3), (499, 4)):
        x = quad.random_quads(rng, nn); y = quad.random_quads(rng, nn)
        want = orc.dot(nn, x, 1, y, 1, T)
        lo, hi = qd.dot_shard_range(nn, T, world, rank, True)
        out = torch.zeros((1, 2), dtype=torch.int64, device="cuda")
        qd.qdot_sharded(nn, to_dev(x[lo:hi]) if hi > lo else torch.zeros((0, 2), dtype=torch.int64, device="cuda"),
                        to_dev(y[lo:hi]) if hi > lo else torch.zeros((0, 2), dtype=torch.int64, device="cuda"), T, out, reference_order=True)
        torch.cuda.synchronize()
        if not (to_host(out).reshape(2) == want).all():
            fails.append(("dot", nn, T))
    # ---- qdot / qnrm2, fast mode: contiguous ranges, all-gather of the 16-byte partials, fixed-order fold on every rank
    qb.set_mode(qb.MODE_FAST)
    nn = 300001
    x = quad.random_quads(rng, nn); y = quad.random_quads(rng, nn)
    lo, hi = qd.dot_shard_range(nn, 1, world, rank, False)
    out = torch.zeros((1, 2), dtype=torch.int64, device="cuda")
    qd.qdot_sharded(nn, to_dev(x[lo:hi]), to_dev(y[lo:hi]), 1, out, reference_order=False)
    torch.cuda.synchronize()
    exact, ratio, klass = orc.exact_dot_check("R", nn, x, nn, y, 1, np.array([[0, 0]]), to_host(out).reshape(1, 2))
    if not (ratio[0] <= 1.0 and klass[0] == 0):
        fails.append(("fast dot contract", float(ratio[0])))
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    if not all((g == out).all() for g in gathered):
        fails.append(("fast dot differs between ranks",))
    qb.set_mode(qb.MODE_REFERENCE)
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    if fails:
        print(f"rank {rank} FAILS: {fails}", flush=True)
    if rank == 0:
        print("mgpu_worker: all passed" if t.item() == 0 else f"mgpu_worker: {int(t.item())} failures", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
